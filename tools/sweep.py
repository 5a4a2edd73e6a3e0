"""Builds tuning variants of librptr_cuda.so (compile-time knobs of the trace kernel) into gpurun_out-independent
paths under build/variants/ and writes run_sweep.sh, which benches each variant on the GPU box:
    python tools/sweep.py NAME:-DX=1,-DY=2 ...   &&   gpurun -- 'bash run_sweep.sh'
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from realtimepathtracingresearchframework_b200 import build as b  # noqa: E402

out_dir = os.path.join(ROOT, "variants")
os.makedirs(out_dir, exist_ok=True)
lines = []
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    lib = os.path.join(out_dir, "librptr_cuda_%s.so" % name)
    cmd = [b.nvcc_path(), "-ccbin", "/usr/bin/g++"] + b.NVCC_FLAGS + [d for d in defs.split(",") if d] + ["-o", lib] + \
          [os.path.join(b.CSRC, f) for f in b.SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(r.stderr[-3000:])
    lines.append("RPTR_CUDA_LIB=variants/librptr_cuda_%s.so timeout 300 python bench.py --steps 2 --warmup 2 %s --no-cpu-baseline "
                 "> gpurun_out/sweep_%s.json 2> gpurun_out/sweep_%s.err || tail -3 gpurun_out/sweep_%s.err" % (name, os.environ.get("SWEEP_ARGS", "--spp 16"), name, name, name))
    print("built", name, defs)
lines.append("python tools/sweep_report.py")
with open(os.path.join(ROOT, "run_sweep.sh"), "w") as f:
    f.write("mkdir -p gpurun_out\n" + "\n".join(lines) + "\n")
