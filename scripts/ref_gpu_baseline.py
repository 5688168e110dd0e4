"""Reference-algorithm GPU baseline as a stand-alone run: bench.gpu_reference on one batch.

    python scripts/ref_gpu_baseline.py [--workload mseg3d_nuscenes|sdseg3d_semantickitti] [--steps 5] [--frames 3]

bench.py reports the same measurement as ``gpu_reference`` on every line; this script exists for longer runs."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mseg3d_nuscenes", choices=["mseg3d_nuscenes", "sdseg3d_semantickitti", "mseg3d_waymo"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--frames", type=int, default=0)
    args = ap.parse_args()
    sys.argv = ["bench.py"]
    import bench
    from lidarseg3d_b200 import synth
    assert torch.cuda.is_available(), "the GPU baseline needs a CUDA device"
    wl = bench.WORKLOADS[args.workload]
    spec = getattr(synth, wl["spec"])
    nf = args.frames or wl["frames_per_gpu"]
    cfg, model = bench.build_model(wl)
    batch = bench.make_batches(wl, spec, 1, nf, 0, n_image_sets=1)[0]
    print(json.dumps(dict(bench.gpu_reference(wl, spec, cfg, model, batch, nf, torch.device("cuda"), steps=args.steps),
                          workload=args.workload)))


if __name__ == "__main__":
    main()
