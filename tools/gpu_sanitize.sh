#!/bin/bash
# compute-sanitizer memcheck of one B=2 UNet evaluation per kernel-variant set (PF_CUDA_GRAPH=0)
tag=${1:-r4w}
out=gpurun_out/$tag
mkdir -p $out
cat > $out/san.py <<'PY'
import torch, sys, os
sys.path.insert(0, '.')
os.environ["PF_CUDA_GRAPH"] = "0"
from bench import sdf_kwargs
from polyffusion_b200.stable_diffusion.model.unet import UNetModel
torch.manual_seed(0)
m = UNetModel(**sdf_kwargs()).eval().cuda()
x = torch.randn(2, 2, 128, 128, device="cuda"); c = torch.randn(2, 1, 512, device="cuda")
t = torch.randint(0, 1000, (2,), device="cuda")
with torch.no_grad():
    y = m(x, t, c)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
PY
run() { name=$1; shift; env PF_CUDA_GRAPH=0 "$@" timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python $out/san.py > $out/sanitizer_memcheck_$name.log 2>&1; echo "memcheck $name rc=$?"; tail -3 $out/sanitizer_memcheck_$name.log; }
run default PF_X=0
run raw_all PF_RAW_SKIP=1 PF_RAW_LN=1
