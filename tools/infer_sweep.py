"""BASELINE.json configs[4]: inference-only encode -> sample -> decode sweep, batch 1..1024 at 256x256 on one B200.
Eval mode (running statistics), eps drawn on the device; latency = median of 5 timed calls after 2 warm-ups (CUDA events)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from face_vae_b200.models import FaceVAE

def main():
    torch.manual_seed(0)
    m = FaceVAE().cuda().eval()
    rows = []
    with torch.no_grad():
        for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
            x = torch.rand((n, 3, 256, 256), device="cuda")
            eps = torch.randn((n, 4096), device="cuda")
            for _ in range(2):
                m(x, True, eps)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); m(x, True, eps); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            rows.append({"batch": n, "latency_ms": ts[2], "images_per_sec": n / ts[2] * 1e3})
            print(f"batch {n:5d}  latency {ts[2]:9.3f} ms  {n / ts[2] * 1e3:10.1f} img/s", flush=True)
            del x, eps
            torch.cuda.empty_cache()
    print(json.dumps({"metric": "inference images/sec at 256x256 (encode->sample->decode, eval mode)", "sweep": rows}))

if __name__ == "__main__":
    main()
