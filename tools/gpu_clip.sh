#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_cabi.py -m gpu -q > gpurun_out/pytest_clip.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_clip.log
timeout 300 python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from ccedit_b200.clip_text import FrozenCLIPEmbedder
from oracle import clip_oracle as co
emb = FrozenCLIPEmbedder(device="cuda")
sd = co.seeded_state_dict({k: tuple(v.shape) for k, v in emb.transformer.state_dict().items()})
emb.transformer.load_state_dict(sd); emb = emb.cuda()
ids = torch.randint(0, 49408, (2, 77), generator=torch.Generator().manual_seed(5))
out = emb(ids.cuda()); ref = co.clip_text_forward(sd, ids)
e = (out.cpu() - ref).abs()
print("clip text: max-norm err %.3e mean err %.3e" % (float(e.max() / ref.abs().max()), float(e.mean() / ref.abs().mean())))
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): emb(ids.cuda())
e1.record(); torch.cuda.synchronize(); print("ms per encode (2 prompts):", e0.elapsed_time(e1) / 10)
PY
