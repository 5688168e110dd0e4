"""N>1 host logic on CPU (gloo, world_size 2): frames shard over ranks with no data-path collective, the timing
reduction is a MAX over ranks, and per-rank batches are disjoint."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.argv = ["bench.py"]
    import bench
    from lidarseg3d_b200 import synth
    spec = dict(synth.KITTI)
    spec.update(beams=4, azimuths=64)
    wl = bench.WORKLOADS["sdseg3d_semantickitti"]
    batches = bench.make_batches(wl, spec, 2, 2, rank)
    sig = float(sum(f.double().sum() for b in batches for f in b["frames"]))
    ms = bench.reduce_max_ms(10.0 + 5.0 * rank, None)
    frames_total = bench.global_frames(steps=3, frames_per_gpu=2, world=world)
    q.put((rank, sig, ms, frames_total, [f.shape[0] for b in batches for f in b["frames"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in procs)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, s0, ms0, n0, _), (r1, s1, ms1, n1, _) = res
    assert s0 != s1                                  # ranks process different frames
    assert ms0 == ms1 == 15.0                        # MAX over ranks
    assert n0 == n1 == 3 * 2 * 2                     # whole-job frame count = steps * frames/GPU * world


def test_config1_cpu_plumbing():
    """BASELINE.json configs[0]: SemanticKITTI-shaped 8 192-point frame, voxelize + forward on the CPU (oracle) with a
    registry-built model's parameters; plumbing only."""
    sys.path.insert(0, ROOT)
    from lidarseg3d_b200.det3d import Config, build_detector
    from oracle import nets as on
    from oracle import voxelize as ov
    cfg = Config.fromfile(os.path.join(ROOT, "configs", "sdseg3d_semantickitti.py"))
    torch.manual_seed(0)
    m = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg).eval()
    rng = np.random.default_rng(0)
    n = 8192
    pts = np.concatenate([rng.uniform(-50, 50, (n, 2)), rng.uniform(-3, 1, (n, 1)), rng.uniform(0, 1, (n, 1))], 1).astype(np.float32)
    vs, rg = cfg.voxel_size, cfg.point_cloud_range
    v, c, nump = ov.points_to_voxel(pts, vs, rg, 5, 300000)
    vv, cc, nn_, nv, p = ov.collate_frames([(v, c, nump, pts)])
    grid = np.round((np.asarray(rg[3:], np.float32) - np.asarray(rg[:3], np.float32)) / np.asarray(vs, np.float32)).astype(np.int64)
    ex = dict(voxels=torch.from_numpy(vv), coordinates=torch.from_numpy(cc), num_points=torch.from_numpy(nn_),
              num_voxels=torch.from_numpy(nv), shape=np.stack([grid]), points=torch.from_numpy(p))
    sd = {k: t.detach() for k, t in m.state_dict().items()}
    with torch.no_grad():
        out = on.segnet_forward(sd, ex, dict(voxel_size=vs, pc_range=rg, reader=dict(type="TransformerVoxelFeatureExtractor",
                                                                                      num_head=4, num_layers=3)))
    assert out.shape == (n, 20) and torch.isfinite(out).all()
    labels = on.predict_labels(out, ex["points"], 1)
    assert labels[0].shape[0] == n
