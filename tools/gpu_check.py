"""Diagnostic run on a GPU box: per-quantity parity errors of the CUDA path against every golden fixture.
Writes gpurun_out/gpu_check.txt.  (Not a test; tests/test_gpu_parity.py asserts the same numbers.)"""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from adt_b200 import testing  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "gpu_check.txt"), "w")


def p(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    out.write(s + "\n")
    out.flush()


p(torch.cuda.get_device_name(0), torch.version.cuda)
prec = "bf16" if "--bf16" in sys.argv else "fp32"
names = [a for a in sys.argv[1:] if not a.startswith("--")] or testing.golden_names()
for name in names:
    try:
        errs = testing.check_golden(name, precision=prec)
        torch.cuda.synchronize()
        for k, v in errs.items():
            flag = "" if v <= testing.tolerance(k) else "   <-- FAIL"
            p(f"{name:12s} {k:72s} {v:.3e}{flag}")
    except Exception:
        p(name, "EXCEPTION")
        p(traceback.format_exc())
out.close()
