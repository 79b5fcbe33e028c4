"""CPU oracle for the Domain-RAG retrieve-then-compose hot path.

TEST INFRASTRUCTURE ONLY. Nothing in the product package (domain_rag_b200/) imports this; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.

Each function restates, in plain numpy / PyTorch fp32 on the CPU, what the reference computes at a
cited file:line. The arithmetic of the reference lives in third-party packages that are NOT in
/root/reference and not installable offline (requirements.txt pins): openai/CLIP @ dcba3cb2,
torchvision 0.22.0, faiss-cpu 1.10.0, diffusers 0.33.1, transformers 4.46.3. Pinning status:

  * stem statistics, style re-rank, DP split, resolution / mask helpers: PINNED - checked against
    golden vectors produced by importing the reference's own functions (oracle/make_golden.py,
    fixtures under tests/golden/).
  * inner-product top-k (faiss), CLIP ViT (openai/CLIP), Flux MMDiT / sampler / Redux blend
    (diffusers): PARITY UNPINNED by the reference (it ships no tests or vectors and the packages
    are absent). They follow the published algorithms and are cross-checked against independent
    local implementations (transformers CLIPVisionModelWithProjection, torchtitan's BFL Flux).
"""
