#!/usr/bin/env python
"""Drop-in entry point for the reference's retrieval/clip100_resnet_style_all_shots.py: same flags
(reference :967-996), same cache and result files (RESULTS_DIR ./retrieval_results, LAMAINPAINT_DIR
../lamainpaint, reference :44-46), served by libdomainrag_b200.so (CLIP ViT + ResNet stem statistics +
inner-product scan x top-k on the B200). Logic: domain_rag_b200/retrieval_cli.py."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from domain_rag_b200.retrieval_cli import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
