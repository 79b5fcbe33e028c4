#!/usr/bin/env python
"""Fill <weights_dir>/text_embeds.pt - the per-prompt text constants of the Redux prior - from the REAL encoders.

The reference instantiates CLIP-L text and T5-XXL in load_model (batch_generate_flux_kshot.py:120-137,
outpainting_updown_sampling_redux.py:505-522) but only ever encodes a handful of fixed strings: "" (every dataset, both
scripts) and the FISH sentence (outpainting...:85-95). This one-shot script encodes exactly those once, with the same
library calls diffusers' FluxPriorReduxPipeline.encode_prompt makes (CLIP: padding="max_length", max_length 77,
pooler_output; T5: padding="max_length", max_length 512, last_hidden_state), so the composition / generation entry
points never need the 9.5 GB T5 on the GPU. It needs the checkpoints on disk (a FLUX.1-dev style directory with
text_encoder/, tokenizer/, text_encoder_2/, tokenizer_2/); there is no download and no synthetic fallback here.

    python scripts/make_text_embeds.py --flux_dir /models/FLUX.1-dev --out ./model/text_embeds.pt [--prompt "extra"]...

Output layout (redux.TextEmbeddingTable.load_file): {"prompts": [[prompt, prompt_2], ...], "t5": bf16 [n,512,4096],
"pooled": bf16 [n,768]}.
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))


def prompt_pairs(extra):
    from domain_rag_b200.hostlogic import _TABLE
    pairs = [("", "")]
    for prm in _TABLE.values():          # compose: prompt=redux_prompt, prompt_2="" (outpainting...:1239-1240)
        if prm.redux_prompt and (prm.redux_prompt, "") not in pairs:
            pairs.append((prm.redux_prompt, ""))
    for p in extra or []:
        if (p, p) not in pairs:
            pairs.append((p, p))
    return pairs


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--flux_dir", required=True, help="directory with text_encoder/, tokenizer/, text_encoder_2/, tokenizer_2/")
    ap.add_argument("--out", default="./model/text_embeds.pt")
    ap.add_argument("--prompt", action="append", help="additional prompt (used as prompt and prompt_2)")
    ap.add_argument("--device", default="cpu")
    args = ap.parse_args(argv)
    need = [os.path.join(args.flux_dir, d) for d in ("text_encoder", "tokenizer", "text_encoder_2", "tokenizer_2")]
    lacking = [d for d in need if not os.path.isdir(d)]
    if lacking:
        print(f"error: missing encoder directories: {lacking}", file=sys.stderr)
        return 2
    import torch
    from transformers import CLIPTextModel, CLIPTokenizer, T5EncoderModel, T5TokenizerFast
    tok1 = CLIPTokenizer.from_pretrained(need[1])
    enc1 = CLIPTextModel.from_pretrained(need[0], torch_dtype=torch.bfloat16).to(args.device).eval()
    tok2 = T5TokenizerFast.from_pretrained(need[3])
    enc2 = T5EncoderModel.from_pretrained(need[2], torch_dtype=torch.bfloat16).to(args.device).eval()
    pairs = prompt_pairs(args.prompt)
    t5s, pooled = [], []
    with torch.no_grad():
        for p1, p2 in pairs:
            ids1 = tok1([p1], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
            pooled.append(enc1(ids1.to(args.device), output_hidden_states=False).pooler_output[0].to("cpu", torch.bfloat16))
            ids2 = tok2([p2], padding="max_length", max_length=512, truncation=True, return_tensors="pt").input_ids
            t5s.append(enc2(ids2.to(args.device), output_hidden_states=False)[0][0].to("cpu", torch.bfloat16))
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    torch.save({"prompts": [list(p) for p in pairs], "t5": torch.stack(t5s), "pooled": torch.stack(pooled)}, args.out)
    print(f"wrote {len(pairs)} prompt pairs to {args.out}: {pairs}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
