#!/usr/bin/env python
"""Benchmark of the Rep-YOLO deployed hot path (fused convs -> Detect decode -> NMS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch 64] [--size 640]

One "step" = one pass of the hot path over one batch of synthetic images (BASELINE.json configs[1]: batch 64 at 640x640,
decode + NMS(conf 0.25, iou 0.45) in the loop, per GPU; N GPUs = N such shards + one NCCL gather of the detections).
Prints ONE JSON line (contract in the task statement): value = images/s with inputs resident in HBM, e2e = the same through
the public API with pinned HOST inputs (H2D + D2H inside the timed region), roofline of the dominant kernel, cpu_baseline.
`--impl reference` times the CPU oracle port of the reference path on the host cores instead (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'fused Rep-YOLO imgs/sec @640 bf16 incl. decode+NMS at 1/2/4/8 B200'
CONF, IOU = 0.25, 0.45
GFLOP_PER_IMAGE_640 = 68.875      # 2*MAC over all executed convs (dense 68.733)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region through NVML (every 5 ms, same counters `nvidia-smi
    --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints; nvidia-smi itself block-buffers a piped stdout)."""
    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.rows, self.halt, self.max_mhz = index, uuid, [], threading.Event(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(('GPU-' + self.uuid).encode())
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
                ids = [v for v in vis.split(',') if v.strip().isdigit()]
                h = nv.nvmlDeviceGetHandleByIndex(int(ids[self.index]) if self.index < len(ids) else self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.halt.is_set():
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), r))
                time.sleep(0.005)
        except Exception:
            self._smi_fallback()

    def _smi_fallback(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self.halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
                c = [v.strip() for v in out.split(',')]
                self.max_mhz = int(c[1])
                bits = sum(b for (_, b), v in zip(self.REASONS, c[2:6]) if v.lower().startswith('active'))
                self.rows.append((int(c[0]), bits))
            except Exception:
                time.sleep(0.05)

    def wait_first(self, timeout=10.0):
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.005)

    def mark(self):
        return len(self.rows)

    def stop(self, first=0):
        self.halt.set()
        rows = self.rows[first:]
        sm = sorted(r[0] for r in rows)
        reasons = sorted({name for name, bit in self.REASONS for r in rows if r[1] & bit})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': reasons, 'samples': len(sm)}


def oracle_step(fz, layers, save, x, torch, O, nms_oracle):
    """the reference's CPU path (oracle port): fused forward + non_max_suppression"""
    _, pred, _ = O.forward_fused(fz, layers, save, x)
    return nms_oracle.non_max_suppression(pred, CONF, IOU)


def run_reference(args):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import nms_oracle
    from oracle import repyolo_oracle as O
    nms_oracle.build()
    layers, save, sd, fz = O.make_model(seed=0, mode=args.init)
    sample = max(1, min(args.ref_sample, args.batch))
    g = torch.Generator().manual_seed(1000)
    x = torch.rand(sample, 3, args.size, args.size, generator=g)
    for _ in range(args.warmup):
        oracle_step(fz, layers, save, x, torch, O, nms_oracle)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(fz, layers, save, x, torch, O, nms_oracle)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    cores = torch.get_num_threads()
    desc = f'{sample} of the {args.batch} images of one step per step (fp32 forward + NMS), {args.steps} steps'
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args),
            'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': desc,
                             'host_cpus': os.cpu_count()},
            'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {'workload': f'Rep-YOLO fused, batch {args.batch} per GPU at {args.size}x{args.size}, Detect decode + NMS '
                        f'(conf {CONF}, iou {IOU}) in-loop; weights: synthetic {args.init} init (seed 0)',
            'batch_per_gpu': args.batch, 'img_size': args.size, 'conf_thres': CONF, 'iou_thres': IOU,
            'l2_policy': 'inputs larger than L2 (fp32 image batch = %.0f MB, uint8 batch = %.0f MB; activations 5 GB per step)' % (
                args.batch * 3 * args.size * args.size * 4 / 1e6, args.batch * 3 * args.size * args.size / 1e6),
            'parallelism': f'batch-sharded dp{args.gpus}, NCCL all-gather of [B,300,6] detections' if args.gpus > 1 else 'single GPU'}


def run_native(args):
    import torch
    import torch.distributed as dist
    import repyolo_b200 as R
    from oracle import repyolo_oracle as O          # weights generator + the cpu_baseline leg only

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)'
    dev = torch.device(f'cuda:{local}')
    torch.cuda.set_device(dev)
    B, S = args.batch, args.size

    layers, save, sd, fz = O.make_model(seed=0, mode=args.init)
    model = R.Model()
    model.load_state_dict(sd, strict=True)
    model.fuse()
    g = torch.Generator().manual_seed(1000 + rank)
    n_bufs = 2
    # e2e ships uint8 NCHW images to the device exactly like the reference's detect.py:73-78 (torch.from_numpy(img).to(device),
    # then .float() / 255 on the device); the native stem fuses that /255.  The kernel-resident `value` leg feeds the same
    # pixels as the fp32 [0,1] tensor Model.forward is specified for.
    host8 = [torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).pin_memory() for _ in range(n_bufs)]
    host = [h.float() / 255.0 for h in host8]
    xdev = [h.to(dev) for h in host]
    eng = model.engine(dev)
    eng.bind(B, S, S)
    launches_per_step = eng.launch_count() + 3    # + NMS: filter, one-kernel sort (cap <= 32768 candidates per image), scan

    def step(x):
        pred, _ = model(x)
        out, counts = R.nms_padded(pred, CONF, IOU)
        if world > 1:
            out, counts = R.gather_detections(out, counts)
        return out, counts

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- kernel-resident throughput (inputs already in HBM) ----
    sampler = ClockSampler(local, str(getattr(torch.cuda.get_device_properties(dev), 'uuid', '') or ''))
    sampler.start()
    for i in range(args.warmup):
        step(xdev[i % n_bufs])
    sync_all()
    sampler.wait_first()
    for i in range(args.warmup):                      # GPU busy again right before the timed region (the wait above idled it)
        step(xdev[i % n_bufs])
    sync_all()
    mark = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out, counts = step(xdev[i % n_bufs])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)

    # ---- end to end through the public API: pinned host batch -> H2D -> forward -> NMS -> D2H of the detections ----
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    staged = [torch.empty((B, 3, S, S), dtype=torch.uint8, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    res_out = torch.empty((B * world, 300, 6), dtype=torch.float32).pin_memory()
    res_cnt = torch.empty((B * world,), dtype=torch.int32).pin_memory()

    def e2e_loop(n):
        for j in range(2):
            freed[j].record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[0])
            staged[0].copy_(host8[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < n:                             # prefetch the next batch while this one computes
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(freed[nxt])
                    staged[nxt].copy_(host8[(i + 1) % n_bufs], non_blocking=True)
                    ready[nxt].record(copy_stream)
            main.wait_event(ready[cur])
            o, c = step(staged[cur])
            freed[cur].record(main)
            res_out.copy_(o, non_blocking=True)
            res_cnt.copy_(c, non_blocking=True)
        main.synchronize()

    e2e_loop(max(2, args.warmup))
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps)
    t1.record()
    sync_all()
    ms_e2e = t0.elapsed_time(t1)
    clocks = sampler.stop(mark)                        # samples taken from the start of the first timed region to the end of the second

    # ---- roofline of the dominant kernel (conv_umma_kernel): per-op CUDA events over K more steps ----
    ops = eng.plan_ir.ops
    eng.set_profiling(True)
    conv_ms, all_ms, per_op = 0.0, 0.0, [0.0] * len(ops)
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        pred, _ = model(xdev[i % n_bufs])
        torch.cuda.synchronize(dev)
        for j, t in enumerate(eng.op_times_ms()):
            per_op[j] += t
    eng.set_profiling(False)
    conv_flops = 0.0
    pk = peaks()
    ridge = pk['tf_sustained'] * 1e12 / (pk['hbm'] * 1e9)          # FLOP per HBM byte at which a conv turns tensor-bound
    cls = {'tensor': [0.0, 0.0, 0.0, 0], 'hbm': [0.0, 0.0, 0.0, 0]}  # [flops, bytes, ms, launches]
    for j, d in enumerate(ops):
        all_ms += per_op[j]
        if d.kind in (2, 11, 12):    # RY_OP_CONV, RY_OP_DETECT (conv_umma_kernel), RY_OP_CONV_CHAIN (conv_chain_kernel)
            conv_ms += per_op[j]
            lvl = eng.plan_ir.tensors[d.in0.tensor].level
            hi, wi = S >> lvl, S >> lvl
            ho, wo = hi // d.stride, wi // d.stride
            fl = 2.0 * d.cout * d.cin * d.ksize * d.ksize * ho * wo * B
            by = 2.0 * B * (hi * wi * d.cin + ho * wo * d.cout / (4 if d.level_idx == 1 else 1)) if d.kind == 2 else B * (2.0 * hi * wi * d.cin + 8.0 * ho * wo * d.cout)
            if d.kind == 12:          # fused chain: all stages' MACs; bytes = input + the outputs that are actually stored
                prev, outs = d.cout, [d.out0, d.out1, d.out2]
                by = 2.0 * B * hi * wi * d.cin + sum(2.0 * B * hi * wi * o.c_len for o in outs if o.tensor >= 0)
                for i in range(d.n_post):
                    fl += 2.0 * prev * d.post_cout[i] * hi * wi * B
                    prev = d.post_cout[i]
            conv_flops += fl
            c = cls['tensor' if fl / by >= ridge else 'hbm']
            c[0] += fl; c[1] += by; c[2] += per_op[j]; c[3] += 1
    conv_ms /= prof_steps
    all_ms /= prof_steps
    n_conv = sum(1 for d in ops if d.kind in (2, 11, 12))
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    detail = {}
    for name, (fl, by, t, n) in cls.items():
        t /= prof_steps
        if n and t > 0:
            detail[name + '_bound_layers'] = {
                'launches': n, 'ms_per_step': t, 'tflops': fl / (t * 1e-3) / 1e12, 'frac_of_bf16_peak': fl / (t * 1e-3) / 1e12 / pk['tf_sustained'],
                'algorithmic_gbytes_per_s': by / (t * 1e-3) / 1e9, 'frac_of_hbm_peak': by / (t * 1e-3) / 1e9 / pk['hbm']}
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
    if os.path.exists(tpath):           # dram__bytes_read+write per conv launch from the committed ncu capture of this workload
        tj = json.load(open(tpath))
        if tj.get('batch') == B and tj.get('size') == S:
            traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')

    tm = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, ms_e2e = tm.tolist()
    value = B * world * args.steps / (ms * 1e-3)
    e2e_v = B * world * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
                'data': 'synthetic', 'config': workload_config(args), 'clocks': clocks,
                'e2e': {'value': e2e_v, 'unit': 'images/s', 'h2d_bytes_per_step': B * 3 * S * S,
                        'input': 'uint8 NCHW from pinned host memory, /255 fused in the stem kernel (reference: detect.py:73-78)',
                        'd2h_bytes_per_step': B * world * (300 * 6 * 4 + 4), 'ms_per_step': ms_e2e / args.steps,
                        'pipeline': 'H2D of batch i+1 overlaps compute of batch i (2 pinned buffers, copy stream)'},
                'gpu_launches': launches_per_step * args.steps,
                'roofline': {'bound': 'tensor', 'kernel': 'conv_umma_kernel (+ conv_chain_kernel: fused 3x3->1x1 chains of the same design)', 'achieved': achieved, 'peak': pk['tf_sustained'],
                             'unit': 'TFLOP/s', 'frac': achieved / pk['tf_sustained'], 'traffic': traffic, 'traffic_source': traffic_src,
                             'peak_source': f"bf16_tflops_sustained of {pk['src']} (kernel timed inside a long step)",
                             'launches_per_step': n_conv, 'avg_launch_ms': conv_ms / max(n_conv, 1),
                             'conv_ms_per_step': conv_ms, 'all_ops_ms_per_step': all_ms,
                             'algorithmic_gflop_per_image': conv_flops / B / 1e9,
                             'algorithmic_bytes_per_launch': sum(c[1] for c in cls.values()) / max(n_conv, 1),
                             'ridge_flop_per_byte': ridge, 'detail': detail},
                'cpu_baseline': None}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import nms_oracle
            nms_oracle.build()
            n = max(1, min(args.ref_sample, B))
            xs = host[0][:n].clone()
            oracle_step(fz, layers, save, xs[:1], torch, O, nms_oracle)
            t0 = time.perf_counter()
            reps = 0
            while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 20):
                oracle_step(fz, layers, save, xs, torch, O, nms_oracle)
                reps += 1
            dt = time.perf_counter() - t0
            line['cpu_baseline'] = {'value': n * reps / dt, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                    'sample': f'{reps} x {n} images of the step batch (fp32 oracle forward + oracle NMS)',
                                    'host_cpus': os.cpu_count()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--size', type=int, default=640)
    ap.add_argument('--init', default='calibrated', choices=['calibrated', 'default'])
    ap.add_argument('--ref-sample', type=int, default=4, help='images per step of the CPU reference arm / cpu_baseline')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'native':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


if __name__ == '__main__':
    main()
