"""Reference-algorithm GPU baseline ("GPU reference-algorithm baseline on the same B200").

Times the oracle's restatement of the reference forward on a CUDA device: spconv-1.x-style sparse convolutions (per kernel
offset index_select -> mm -> index_add_, separate BatchNorm / ReLU), chunked-distance 3-NN, the heads with their per-frame
python loops, HRNet / FCN through cuDNN in fp32 with TF32 off - what the reference does on a GPU, minus real spconv (not
installable here).  This is the denominator of the north-star ">= 10x the reference spconv-GPU forward" target; it is a
reported baseline and no part of the product path.

    python scripts/ref_gpu_baseline.py [--workload mseg3d_nuscenes|sdseg3d_semantickitti] [--steps 5] [--frames 3]

Prints one JSON line.  NOT YET RUN ON A GPU (written in round 1 after the GPU budget was spent; its CPU equivalence with
the numpy oracle is tested in tests/test_oracle_torch_backend.py).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mseg3d_nuscenes", choices=["mseg3d_nuscenes", "sdseg3d_semantickitti"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--frames", type=int, default=0)
    args = ap.parse_args()
    sys.argv = ["bench.py"]
    import bench
    from lidarseg3d_b200 import synth
    from oracle import nets as on
    from oracle import torch_backend as tb
    from oracle import voxelize as ov
    assert torch.cuda.is_available(), "the GPU baseline needs a CUDA device"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda")
    wl = bench.WORKLOADS[args.workload]
    spec = getattr(synth, wl["spec"])
    nf = args.frames or wl["frames_per_gpu"]
    cfg, model = bench.build_model(wl)
    sd = {k: v.detach().to(dev) for k, v in model.state_dict().items()}
    batch = bench.make_batches(wl, spec, 1, nf, 0)[0]
    frames = [f.numpy() for f in batch["frames"]]
    # the reference voxelizes on the loader's CPU (numba); its result is an input of the GPU forward
    vox = [ov.points_to_voxel(f, spec["voxel_size"], spec["pc_range"], 5, 300000) for f in frames]
    v, c, n, nv, pts = ov.collate_frames([(a, b, cc, f) for (a, b, cc), f in zip(vox, frames)])
    ex = dict(voxels=torch.from_numpy(v).to(dev), coordinates=torch.from_numpy(c).to(dev), num_points=torch.from_numpy(n).to(dev),
              num_voxels=torch.from_numpy(nv), shape=np.stack([synth.grid_shape(spec)] * nf), points=torch.from_numpy(pts).to(dev))
    if wl["cam"]:
        ex["points_cuv"] = batch["cuv"].to(dev)
        ex["images"] = torch.from_numpy(on.image_input_transform(batch["images_u8"].numpy(), synth.IMG_MEAN, synth.IMG_STD)).to(dev)
        ocfg = dict(voxel_size=spec["voxel_size"], pc_range=spec["pc_range"], hrnet_extra=cfg.model.img_backbone.extra, nhead=4,
                    nlayer=6, num_convs=2)
        fwd = lambda: on.mseg3d_forward(sd, ex, ocfg, backend=tb)
    else:
        ocfg = dict(voxel_size=spec["voxel_size"], pc_range=spec["pc_range"],
                    reader=dict(type="TransformerVoxelFeatureExtractor", num_head=4, num_layers=3))
        fwd = lambda: on.segnet_forward(sd, ex, ocfg, backend=tb)
    with torch.no_grad():
        for _ in range(args.warmup):
            fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = fwd()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps(dict(metric="reference_algorithm_gpu_forward_frames_per_sec", value=nf / (ms / 1e3), unit="frames/s",
                          ms_per_step=ms, frames_per_step=nf, workload=args.workload, steps=args.steps,
                          dtype="fp32 (TF32 off)", points=int(pts.shape[0]), logits_shape=list(out.shape),
                          note="oracle restatement of the reference forward on the GPU: spconv-1.x-style per-offset "
                               "gather/mm/scatter-add, separate BN/ReLU, cuDNN fp32 camera branch; voxelization excluded (the "
                               "reference does it in the CPU loader)")))


if __name__ == "__main__":
    main()
