"""Development helper: run every conv parity case in its own process (a trapped kernel poisons the CUDA context)."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(idx):
    import torch  # noqa: F401
    from test_conv_gpu import CASES, run_conv_case

    lib = C.CDLL(os.path.join(ROOT, "dafne_b200", "libdafne_b200.so"))
    lib.dafne_last_error.restype = C.c_char_p
    case = CASES[idx]
    try:
        err = run_conv_case(lib, *case)
        print(f"PASS {case[0]} max_err={err:.4g}", flush=True)
    except AssertionError as e:
        print(f"FAIL {case[0]}: {e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(int(sys.argv[1]))
    else:
        from test_conv_gpu import CASES

        for i in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=120)
                out = (r.stdout + r.stderr).strip().splitlines()
                tail = [l for l in out if l.startswith(("PASS", "FAIL", "dafne"))] or out[-6:]
                print(f"[{i}] rc={r.returncode} " + " | ".join(tail[-6:]), flush=True)
            except subprocess.TimeoutExpired:
                print(f"[{i}] TIMEOUT {CASES[i][0]}", flush=True)
