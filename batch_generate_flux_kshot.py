#!/usr/bin/env python
"""Drop-in entry point for the reference's batch_generate_flux_kshot.py: same flags and output tree, served by
libdomainrag_b200.so on the B200. Logic: domain_rag_b200/generate_cli.py."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))

from domain_rag_b200.generate_cli import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
