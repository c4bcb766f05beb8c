"""Measurement tooling: build named variants of libfmc_b200.so with extra -D flags into build_variants/
(git-ignored, shipped to the GPU box).   python scripts/build_variants.py name:-DA=1,-DB=2 ...
Run them with scripts/gpu_variants.sh name ...  (FMC_LIB_PATH selects the library)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_monte_carlo_b200 import build as b
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(root, "build_variants"), exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out = os.path.join(root, "build_variants", f"libfmc_{name}.so")
    cmd = [b.nvcc_path()] + b.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-Xptxas", "-v", "-o", out, b.SRC]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    log = p.communicate()[0]
    keep = [l for l in log.splitlines() if "error" in l or ("sim_memo_kernelILb0" in l)]
    idx = [i for i, l in enumerate(log.splitlines()) if "Compiling entry function" in l and "sim_memo_kernelILb0" in l]
    lines = log.splitlines()
    info = " | ".join(lines[i + 2].strip() + " " + lines[i + 3].strip() for i in idx[:1]) if idx else ""
    print(name, "rc", p.returncode, info)
    if p.returncode:
        print(log[-2000:])
