"""bench.py --workload {predlift3m, lift10m, raht30m}: the other BASELINE.json
configurations (configs[2]-[4]).  Same JSON contract as the default workload
(configs[1], bench.py); single process per GPU, slices of a frame are the
independent work units (one C-ABI call / lane each)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))

WORKLOADS = {
    # name: (points, max points per slice, description, algorithmic bytes per point)
    "predlift3m": (3_000_000, 1_000_000,
                   "configs[2]: octree-predlift lossless-geom nearlossless-attrs, 3M-point dense surface "
                   "cloud in 3 slices, level-of-detail build (12 LoDs, distance subsampling, k=3, "
                   "intra-LoD search, blended weights)", 104),
    "lift10m": (10_000_000, 1_000_000,
                "configs[3]: octree-liftt lossy-geom lossy-attrs, 10M-point cloud in 10 slices, 3 LoD "
                "levels, RGB: LoD build + weights + forward lifting + quantisation + reconstruction", 104 + 96 + 144 + 68),
    "raht30m": (30_000_000, 1_100_000,
                "configs[4]: octree-raht intra, one 30M-point frame per GPU in 28 slices, RGB + "
                "reflectance in one pass per slice", 80),
}


def lod_params(pb, levels, predicting):
    lp = pb.LodParams()
    lp.num_detail_levels, lp.lod_decimation_type, lp.dist2 = levels, 0, 0
    lp.num_pred_nearest_neighbours, lp.inter_lod_search_range = 3, 1100000
    lp.intra_lod_search_range = 1100000 if predicting else 0
    lp.intra_lod_prediction_skip_layers = 0 if predicting else 0x7fffffff
    lp.prediction_with_distribution, lp.pred_weight_blending = 1, 1 if predicting else 0
    for i in range(3):
        lp.lod_neigh_bias[i] = 1
    for i in range(32):
        lp.lod_sampling_period[i] = 4
    return lp


def run(args, bench):
    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor

    sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
    import pcc_attr_b200 as pb
    from pcc_attr_b200.synth import cloud_terrain, morton_slices, texture

    name = args.workload
    npts, per_slice, desc, alg_bytes = WORKLOADS[name]
    if args.points:
        npts = args.points
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bench.bind_to_gpu_numa_node(torch, local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    pb.lib()
    pb.set_device(local)
    params, qpset = bench.make_pods(pb)
    if distributed:
        if rank != 0:
            params, qpset = pb.RahtParams(), pb.QpSet()
        raw = bench.broadcast_pods(bytes(params) + bytes(qpset), dist, dev)
        params = pb.RahtParams.from_buffer_copy(raw[:C.sizeof(pb.RahtParams)])
        qpset = pb.QpSet.from_buffer_copy(raw[C.sizeof(pb.RahtParams):])

    xyz, rgb = cloud_terrain(npts, seed=7 + rank)
    rgb = texture(rgb, bench.TEXTURE_RGB, 100 + rank)
    refl = texture(((rgb[:, :1] * 2 + rgb[:, 1:2]) // 3).astype(np.int32), bench.TEXTURE_REFL, 200 + rank)
    xyz, (rgb, refl), offs = morton_slices(xyz, [rgb, refl], per_slice)
    ns = len(offs) - 1

    def pinned(a):  # the host-pointer entries stage from / to page-locked memory
        t = torch.empty(a.shape, dtype=torch.int32, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy(), t

    keep_alive = []
    if name != "raht30m":
        xyz, t0_ = pinned(np.ascontiguousarray(xyz, dtype=np.int32))
        rgb, t1_ = pinned(np.ascontiguousarray(rgb, dtype=np.int32))
        keep_alive += [t0_, t1_]
    pool = ThreadPoolExecutor(max_workers=min(ns, 32))
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def run_jobs(jobs):
        for f in [pool.submit(j) for j in jobs]:
            f.result()

    def sl(a, s):
        return a[offs[s]:offs[s + 1]]

    h2d = d2h = 0
    host_jobs = None  # set where the end-to-end leg differs from the device-resident one
    if name == "predlift3m":
        lp = lod_params(pb, 12, True)

        def slice_job(s):
            h = C.c_void_p()
            x = sl(xyz, s)
            pb._check(pb.lib().pccb200_lod_create(C.byref(lp), pb._p(x, C.c_int32), C.c_int32(len(x)), C.byref(h)))
            pb.lib().pccb200_lod_destroy(h)

        jobs = [lambda s=s: slice_job(s) for s in range(ns)]
        h2d, d2h = xyz.nbytes, 0
        resident = "positions are staged by the call (host-pointer C ABI, pinned); the levels of detail stay on the device (handle)"
    elif name == "lift10m":
        lp = lod_params(pb, 3, False)
        lq = pb.QpSet()
        lq.num_layers, lq.max_qp, lq.fixed_point_qp_offset = 1, 51, 24
        lq.layers[0][0], lq.layers[0][1] = bench.QP, 0
        out, t2_ = pinned(rgb)
        vals, t3_ = pinned(rgb)
        keep_alive += [t2_, t3_]
        lcp = np.zeros((ns, 32), dtype=np.int8)
        so = np.ascontiguousarray(offs, dtype=np.int64)

        def all_slices():
            np.copyto(out, rgb)
            pb._check(pb.lib().pccb200_attr_lift_encode_slices(
                C.byref(lp), C.byref(lq), C.c_int32(1), None, pb._p(xyz, C.c_int32), pb._p(out, C.c_int32),
                C.c_int32(3), C.c_int32(8), pb._p(so, C.c_int64), C.c_int32(ns), pb._p(vals, C.c_int32),
                pb._p(lcp, C.c_int8)))

        dxyz = torch.from_numpy(xyz).to(dev)
        drgb0 = torch.from_numpy(rgb).to(dev)
        drgb, dvals = torch.empty_like(drgb0), torch.empty_like(drgb0)
        lcp_d = np.zeros((ns, 32), dtype=np.int8)

        def dev_slices():  # device-resident inputs and outputs, coded in place
            pb.attr_lift_slices_dev(True, lp, lq, 1, dxyz.data_ptr(), drgb.data_ptr(), 3, so,
                                    dvals.data_ptr(), lcp_d)

        jobs = [dev_slices]
        host_jobs = [all_slices]
        h2d, d2h = xyz.nbytes + rgb.nbytes, 2 * rgb.nbytes
        resident = ("value: device-resident inputs and outputs (pccb200_attr_lift_encode_slices_dev); "
                    "e2e: host-pointer C ABI with pinned host buffers (H2D / D2H inside the timed region)")
    else:
        dxyz = torch.from_numpy(xyz).to(dev)
        drgb0, drefl0 = torch.from_numpy(rgb).to(dev), torch.from_numpy(refl).to(dev)
        drgb, drefl = torch.empty_like(drgb0), torch.empty_like(drefl0)
        crgb = [torch.empty((3, int(offs[s + 1] - offs[s])), dtype=torch.int32, device=dev) for s in range(ns)]
        crefl = [torch.empty((1, int(offs[s + 1] - offs[s])), dtype=torch.int32, device=dev) for s in range(ns)]

        def slice_job(s):  # (one slice alone: the per-phase profile leg)
            o, n = int(offs[s]), int(offs[s + 1] - offs[s])
            pb.attr_raht_encode_multi_dev(
                params, [qpset, qpset], dxyz.data_ptr() + 12 * o,
                [drgb.data_ptr() + 12 * o, drefl.data_ptr() + 4 * o],
                [crgb[s].data_ptr(), crefl[s].data_ptr()], n, [3, 1])

        def all_slices():  # the slices of the frame in ONE batch call (coded in gangs)
            pb.attr_raht_multi_batch_dev(
                True, params, [qpset, qpset], [dxyz.data_ptr() + 12 * int(offs[s]) for s in range(ns)],
                [[drgb.data_ptr() + 12 * int(offs[s]), drefl.data_ptr() + 4 * int(offs[s])] for s in range(ns)],
                [[crgb[s].data_ptr(), crefl[s].data_ptr()] for s in range(ns)],
                [int(offs[s + 1] - offs[s]) for s in range(ns)], [3, 1])

        jobs = [all_slices]
        one_slice = lambda: slice_job(0)
        resident = "device-resident inputs and outputs"

    def prepare():
        flush.fill_(1)
        if name == "raht30m":
            drgb.copy_(drgb0)
            drefl.copy_(drefl0)
        elif name == "lift10m":
            drgb.copy_(drgb0)
        torch.cuda.synchronize()

    def step():
        prepare()
        pb.time_begin()
        t0 = time.perf_counter()
        run_jobs(jobs)
        wall = time.perf_counter() - t0
        return pb.time_end(), wall

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    sampler = bench.ClockSampler(local)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    l0 = pb.kernel_launch_count()
    res = [step() for _ in range(args.steps)]
    launches = pb.kernel_launch_count() - l0
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms, wall_s = sum(r[0] for r in res), sum(r[1] for r in res)
    if host_jobs:  # end to end through the host-pointer entry
        for _ in range(2):
            run_jobs(host_jobs)
        wall_s = 0.0
        for _ in range(args.steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run_jobs(host_jobs)
            wall_s += time.perf_counter() - t0
        same = bool(np.array_equal(vals, dvals.cpu().numpy()) and np.array_equal(out, drgb.cpu().numpy()))
        if not same:
            raise SystemExit("bench: device-pointer and host-pointer lifting results differ")

    # one slice alone, per-phase device times
    pb.profile_reset()
    pb.profile_enable(True)
    prepare()
    if name == "lift10m":
        so1 = np.ascontiguousarray(offs[:2], dtype=np.int64)
        o1, v1, l1 = rgb[:offs[1]].copy(), np.empty_like(rgb[:offs[1]]), np.zeros((1, 32), dtype=np.int8)
        pb._check(pb.lib().pccb200_attr_lift_encode_slices(
            C.byref(lp), C.byref(lq), C.c_int32(1), None, pb._p(xyz, C.c_int32), pb._p(o1, C.c_int32),
            C.c_int32(3), C.c_int32(8), pb._p(so1, C.c_int64), C.c_int32(1), pb._p(v1, C.c_int32),
            pb._p(l1, C.c_int8)))
    elif name == "raht30m":
        one_slice()
    else:
        jobs[0]()
    pb.profile_enable(False)
    prof = pb.profile_read()
    slice_ms = sum(v[0] for v in prof.values())
    n0 = int(offs[1] - offs[0])

    per_rank = [dev_ms / args.steps]
    if distributed:
        g = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(g, torch.tensor([dev_ms, wall_s], dtype=torch.float64, device=dev))
        per_rank = [float(x[0]) / args.steps for x in g]
        dev_ms, wall_s = max(float(x[0]) for x in g), max(float(x[1]) for x in g)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        achieved = alg_bytes * n0 / (slice_ms * 1e-3) / 1e9
        line = {
            "metric": bench.METRIC if name == "raht30m" else
            "attribute-transform Mpoints/s (" + ("level-of-detail build" if name == "predlift3m" else "lifting transform") + ")",
            "value": world * npts * args.steps / (dev_ms * 1e-3) / 1e6, "unit": "Mpoints/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic",
            "config": {"workload": desc, "points_per_frame": npts, "slices": ns,
                       "points_per_slice": n0, "inputs": resident,
                       "l2": "512 MiB written between steps (excluded from timing) to flush L2"},
            "e2e": {"value": world * npts * args.steps / wall_s / 1e6, "unit": "Mpoints/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * wall_s / args.steps,
                    "note": "wall clock around the C-ABI calls of a step"},
            "per_rank_ms_per_step": {"min": min(per_rank), "median": float(np.median(per_rank)),
                                     "max": max(per_rank)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "all kernels of one slice, timed alone (per-phase CUDA events)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_slice": alg_bytes * n0, "kernel_ms_per_slice": slice_ms},
            "phase_ms_one_slice_alone": {k: v[0] for k, v in prof.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(name, bench, sl(xyz, 0), sl(rgb, 0), sl(refl, 0), n0)
        bench.emit_json(line)
    pool.shutdown()
    if distributed:
        dist.destroy_process_group()


def cpu_slice_seconds(name, bench, xyz, rgb, refl):
    """one slice through the compiled reference (oracle/_ref) on the calling
    thread; no entropy coding in any of the workloads -> (seconds, what, kind)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pcc_testlib as tl

    kind = "reference" if tl.ref_available() else "port"
    if name == "raht30m":
        run, kind = bench.load_cpu_impl()
        secs = bench.cpu_frame_seconds(run, tl.make_params(search_range=bench.SEARCH_RANGE),
                                       tl.make_qpset(qp=bench.QP, chroma_offset=bench.CHROMA_OFFSET),
                                       (xyz, rgb, refl))
        return secs, "RAHT of RGB and of reflectance", kind
    predicting = name == "predlift3m"
    lp = tl.make_lod_params(levels=12 if predicting else 3, intra_range=1100000 if predicting else 0,
                            skip_layers=0 if predicting else 0x7fffffff, blending=1 if predicting else 0)
    build = tl.ref_lod_build if kind == "reference" else tl.oracle_lod_build
    t0 = time.perf_counter()
    preds, idx, npl = build(lp, xyz)
    what = "AttributeLods::generate"
    if not predicting:
        qw = (tl.ref_quant_weights if kind == "reference" else tl.oracle_quant_weights)(preds)
        lift = tl.ref_lift if kind == "reference" else tl.oracle_lift
        a = (rgb[idx].astype(np.int64) << 8)
        c = lift(True, preds, qw, npl, a)
        rec, vals = tl.oracle_lift_quant(True, tl.make_qpset(qp=bench.QP, chroma_offset=0,
                                                             fixed_point_qp_offset=24), qw, npl, c,
                                         lcp=np.zeros(33, dtype=np.int8))
        lift(False, preds, qw, npl, rec)
        what += (" + quantisation weights + PCCLiftPredict/Update forward and inverse (reference), "
                 "quantisation (oracle port)")
    return time.perf_counter() - t0, what, kind


def cpu_baseline(name, bench, xyz, rgb, refl, n0):
    secs, what, kind = cpu_slice_seconds(name, bench, xyz, rgb, refl)
    return {"value": n0 / secs / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": kind,
            "sample": f"one slice ({n0} points) on one host core: {what}, {secs:.2f} s; "
                      f"host: {bench.host_cpu_model()}"}


def run_reference(args, bench):
    """--impl reference: the reference's CPU implementation of the workload's
    path, one slice per host thread on the physical cores (bounded sample)."""
    import threading

    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
    from pcc_attr_b200.synth import cloud_terrain, morton_slices, texture

    name = args.workload
    npts, per_slice, desc, _ = WORKLOADS[name]
    sample = min(npts, 4 * per_slice)
    xyz, rgb = cloud_terrain(sample, seed=7)
    rgb = texture(rgb, bench.TEXTURE_RGB, 100)
    refl = texture(((rgb[:, :1] * 2 + rgb[:, 1:2]) // 3).astype(np.int32), bench.TEXTURE_REFL, 200)
    xyz, (rgb, refl), offs = morton_slices(xyz, [rgb, refl], per_slice)
    ns = len(offs) - 1
    cores = bench.physical_cores()
    info = {}

    def work(i):
        s = i % ns
        a, b = offs[s], offs[s + 1]
        info["r"] = cpu_slice_seconds(name, bench, xyz[a:b], rgb[a:b], refl[a:b])

    def one_step():
        ts = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    total = sum(one_step() for _ in range(args.steps))
    n0 = int(offs[1] - offs[0])
    value = cores * n0 * args.steps / total / 1e6
    _, what, kind = info["r"]
    bench.emit_json({
        "impl": "reference", "metric": bench.METRIC if name == "raht30m" else
        "attribute-transform Mpoints/s (" + ("level-of-detail build" if name == "predlift3m" else "lifting transform") + ")",
        "value": value, "unit": "Mpoints/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": desc},
        "cpu_baseline": {"value": value, "unit": "Mpoints/s", "cores": cores, "kind": kind,
                         "sample": f"each step: one slice ({n0} points) per host thread, {cores} threads: "
                                   f"{what}; host: {bench.host_cpu_model()}"},
        "e2e": {"value": value, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
