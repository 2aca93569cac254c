"""Bring-up diagnostics: run every parity case on the GPU, never stop at the first failure, and
write a full error table to gpurun_out/check.log (used while developing; pytest -m gpu is the gate)."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle.cases import LAYER_CFGS, MODEL_CFGS, MODULE_CASES, NMS_CASES  # noqa: E402
from tests import parity  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "check.log"), "w")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def guarded(label, fn):
    t = time.time()
    try:
        r = fn()
        torch.cuda.synchronize()
        return r
    except Exception:
        say(f"EXC {label}: {traceback.format_exc()[-1500:]}")
        try:
            torch.cuda.synchronize()
        except Exception as e:  # sticky CUDA error: nothing after this is meaningful
            say("FATAL sticky CUDA error:", e)
            LOG.close()
            os._exit(3)
        return None
    finally:
        pass


def main():
    say("device", torch.cuda.get_device_name(0), "torch", torch.__version__)
    only = sys.argv[1:] or ["modules", "models", "nms"]
    bad = 0
    for dtype in (torch.float32, torch.bfloat16):
        tol = parity.TOL[dtype]
        if "modules" in only:
            for case in MODULE_CASES:
                r = guarded(f"module {case[0]} {dtype}", lambda: parity.run_module_case(case, dtype))
                if r is None:
                    bad += 1
                    continue
                worst = max(v[0] for v in r.values())
                flag = "ok " if worst <= tol else "BAD"
                bad += flag == "BAD"
                say(f"{flag} module {case[0]:24s} {str(dtype)[6:]:9s} " + " ".join(f"{k}={v[0]:.2e}/{v[1]:.2e}" for k, v in r.items()))
        if "models" in only:
            for cfg in MODEL_CFGS:
                r = guarded(f"model {cfg} {dtype}", lambda: parity.run_model_case(cfg, dtype, layers=cfg in LAYER_CFGS))
                if r is None:
                    bad += 1
                    continue
                worst = max(v[0] for v in r.values())
                flag = "ok " if worst <= tol else "BAD"
                bad += flag == "BAD"
                say(f"{flag} model {cfg:32s} {str(dtype)[6:]:9s} worst={worst:.2e}")
                for k, v in r.items():
                    say(f"      {k:34s} max_rel={v[0]:.2e} l2_rel={v[1]:.2e}")
    if "nms" in only:
        for ci in range(len(NMS_CASES)):
            r = guarded(f"nms {NMS_CASES[ci][0]}", lambda: parity.run_nms_case(ci))
            if r is None:
                bad += 1
                continue
            ok = all(x[0] for x in r)
            bad += not ok
            say(f"{'ok ' if ok else 'BAD'} nms {NMS_CASES[ci][0]:20s} {r}")
    say("TOTAL BAD", bad)
    return 0


if __name__ == "__main__":
    sys.exit(main())
