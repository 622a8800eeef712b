#!/usr/bin/env python
"""Benchmark of the FFC head hot path (BASELINE.json metric: head fwd+bwd samples/s, % of bf16 tensor peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4] [--impl ours|reference]

A *step* is one reference ``FFC.forward`` over one synthetic batch, embeddings in: a rollback head pass
(probe=x, gallery=y) plus a commit head pass (probe=y, gallery=x) -- ffc.py:264-267 -- i.e. 2*B probe rows
("samples") per GPU, each through LRU bookkeeping, enqueue scatter, the fused margin-softmax sweep (loss and
dEmb, both add_margin terms) and, on the rollback pass, the restore.  Workloads:
  c3 (default)  B=1024 rows/GPU, 1,048,576 identities, queue 1,048,576 (column-sharded over N GPUs), D=512, Arc
  c2            B=512, 100k identities, queue 65,536, D=512, Arc (single GPU)
  c4            B=512 rows/GPU, 10,000,000 identities, queue 1,048,576 (LRU-managed: evictions, unknown labels), D=512, Arc
`value` is device-timed with the inputs resident in HBM; `e2e` is the same metric through the public
``FFCHead.head`` API from pinned host buffers (H2D of embeddings+labels and a D2H read of the loss inside the
timed region).  `--impl reference` times the reference algorithm's CPU port (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'very-large-scale-face-recognition_b200'))

WORKLOADS = {
    'c3': dict(name='C3: FFC head only, 512-d embeddings, batch 1024/GPU, 1M identities, queue 1M', B=1024, N=1 << 20, Q=1 << 20, D=512,
               loss_type='Arc', margin=0.5, scale=32.0),
    'c2': dict(name='C2: FFC head, batch 512, 100k identities, queue 64k, D=512', B=512, N=100000, Q=65536, D=512,
               loss_type='Arc', margin=0.5, scale=32.0),
    # head of C4 (the backbone is outside the path): ten identities per queue slot, so instance rows mostly miss (LRU eviction path)
    # and most probe labels of the instance half are unknown (hard-negative rows).  Not the default; not measured in round 1.
    'c4': dict(name='C4 head: batch 512/GPU, 10M identities, LRU-managed queue 1M, D=512', B=512, N=10_000_000, Q=1 << 20, D=512,
               loss_type='Arc', margin=0.5, scale=32.0),
}
CPU_SAMPLE = dict(B=256, Q=32768)   # bounded CPU sample: rows per pass and queue slice; scaled linearly in Q


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(burst=d['bf16_tflops'], sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']), hbm=d['hbm_gbs'], src='measured')
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src='fallback')


def sweep_traffic(w, world):
    """DRAM bytes of one main-sweep launch from the committed ncu --set full capture (C3 at one GPU only)."""
    p = os.path.join(ROOT, 'profiles', 'r1_sweep_traffic.json')
    if world == 1 and w is WORKLOADS['c3'] and os.path.isfile(p):
        return json.load(open(p))['dram_bytes_per_launch']
    return None


def config_of(w, world):
    """The `config` object of the JSON line: the same for our arm and the reference arm."""
    return dict(workload=w['name'], rows_per_pass_per_gpu=w['B'], passes_per_step=2, identities=w['N'], queue=w['Q'], feat_dim=w['D'],
                loss=w['loss_type'], margin=w['margin'], scale=w['scale'], sharding=('none' if world == 1 else f'queue columns /{world}'),
                l2_policy='working set (bf16 queue %.0f MB per rank) exceeds the 126 MB L2' % (w['Q'] // world * w['D'] * 2 / 1e6))


def make_batches(w, n_batches, seed, rank=0, world=1):
    """id half = chunks of a seeded permutation (same ids in x and y), instance halves iid uniform."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed + 7919 * rank)
    B, N, D = w['B'], w['N'], w['D']
    h = B // 2
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(seed))
    out = []
    for s in range(n_batches):
        base = ((s * world + rank) * h) % max(1, N - h)
        ids = perm[base:base + h]
        xl = torch.cat([ids, torch.randint(0, N, (B - h,), generator=g)])
        yl = torch.cat([ids, torch.randint(0, N, (B - h,), generator=g)])
        x = F.normalize(torch.randn(B, D, generator=g))
        y = F.normalize(torch.randn(B, D, generator=g))
        out.append((x, y, xl, yl))
    return out


# --------------------------------------------------------------------------------------------------
# CPU baseline: the reference algorithm's port (oracle/head_ref.py) on the host cores, bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_reference(w, steps, warmup):
    import torch
    from oracle.head_ref import HeadOracle
    Bs, Qs = min(CPU_SAMPLE['B'], w['B']), min(CPU_SAMPLE['Q'], w['Q'])
    ws = dict(w, B=Bs, N=max(Qs, int(w['N'] * Qs / w['Q'])), Q=Qs)
    o = HeadOracle(w['D'], Qs, w['scale'], w['loss_type'], w['margin'], dtype=torch.float32)
    o.lru.restore([(i, i) for i in range(Qs)])
    batches = make_batches(ws, 4, seed=1234)
    times = []
    for s in range(warmup + steps):
        x, y, xl, yl = batches[s % len(batches)]
        x = x.clone().requires_grad_(True)
        y = y.clone().requires_grad_(True)
        t0 = time.perf_counter()
        loss = o.forward(x, y, xl.tolist(), yl.tolist())
        loss.backward()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    raw = 2 * Bs / t                               # samples/s at the sample's queue size
    scaled = raw * Qs / w['Q']                     # per-sample cost is linear in Q
    sample = (f'oracle port (torch fp32, materialised B x Q logits, argsort top-k, autograd backward) on B={Bs} rows/pass, '
              f'queue {Qs}, D={w["D"]}: {t * 1e3:.0f} ms/step = {raw:.0f} samples/s, scaled by {Qs}/{w["Q"]} to the full queue')
    return dict(value=scaled, unit='samples/s', cores=torch.get_num_threads(), kind='port', sample=sample), t


def run_reference(args, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb, t = cpu_reference(w, max(1, args.steps), max(0, args.warmup))
    line = dict(metric='ffc_head_fwd_bwd_samples_per_s', value=cb['value'], unit='samples/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=config_of(w, max(1, args.gpus)), cpu_baseline=cb,
                e2e=dict(value=cb['value'], unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------------
class Clocks:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx, pw = [], set(), None, []
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                pw.append(float(r[3]))
                for n, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            sm.sort()
            out = dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm), power_w_max=max(pw) if pw else None)
        return out


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import ffc_b200
    from ffc_b200 import _capi

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _capi.lib()
    B, N, Q, D = w['B'], w['N'], w['Q'], w['D']

    if world == 1:
        head = ffc_b200.FFCHead(D, Q, w['scale'], w['loss_type'], w['margin'], precision='bf16', max_batch=B, device=dev)
        head._ensure()
        # steady state: LRU full: ids 0..Q-1 resident, slot i <- id i, recency = id order
        head.lru.restore_arrays(torch.arange(Q, dtype=torch.int64), torch.arange(Q, dtype=torch.int32))

        def step(x, y, xl, yl):          # rollback pass (probe x, gallery y) + commit pass (probe y, gallery x): loss and both dEmb
            return head.forward_pair(x, y, y, x, xl, yl)[0]

        def step_api(x, y, xl, yl):      # public API (FFC.forward with embeddings in), gradients requested
            return head(x.requires_grad_(True), y.requires_grad_(True), xl, yl)
        prefetch = None                  # one GPU: host labels already run ahead on the bookkeeping stream
    else:
        from ffc_b200.dist import ShardedFFCHead
        head = ShardedFFCHead(D, Q, w['scale'], w['loss_type'], w['margin'], max_batch=B, device=dev)
        head.prefill_identity(N)

        def step(x, y, xl, yl):
            return head.forward_pair(x, y, xl, yl)[0]
        step_api = step
        # the sharded head takes the NEXT step's labels (CPU tensors, main.py:59-60) as soon as the loader has them:
        # their all-gather and the rollback pass's LRU bookkeeping then run under the current step's sweeps
        prefetch = None if os.environ.get('FFC_BENCH_NO_PREFETCH') else head.prefetch

    # one distinct batch per step: re-feeding an embedding that is already in the queue makes the target cosine
    # exactly 1, the reference's Arc NaN hazard
    n_b = args.steps + args.warmup
    host = make_batches(w, n_b, seed=1234, rank=rank, world=world)
    pinned = [tuple(t.pin_memory() for t in b) for b in host]
    # sharded head: labels stay on the host (pinned), as in the reference's loop; one GPU: device-resident
    devb = [(x.to(dev), y.to(dev), xl.pin_memory() if world > 1 else xl.to(dev), yl.pin_memory() if world > 1 else yl.to(dev)) for x, y, xl, yl in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    clocks = Clocks(local) if rank == 0 else None
    for s in range(args.warmup):
        step(*devb[s % n_b])
        if prefetch:
            prefetch(*devb[(s + 1) % n_b][2:])
    barrier()
    if getattr(head, '_timing', None):
        head._timing.clear()
    head.set_timing(True)
    l0 = lib.ffc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        loss = step(*devb[(args.warmup + s) % n_b])
        if prefetch and s + 1 < args.steps:
            prefetch(*devb[(args.warmup + s + 1) % n_b][2:])
    e1.record()
    barrier()
    launches = lib.ffc_launch_count() - l0
    ms = e0.elapsed_time(e1)
    sweep_ms, sweep_n = head.get_timing()
    head.set_timing(False)
    clk = clocks.stop() if clocks else None
    loss_val = float(loss)

    # ---- end to end through the public API from pinned host buffers (e2e) ----
    pinned = pinned[::-1]     # other embeddings than the ones just enqueued
    phase_ms = head.phase_times() if getattr(head, '_timing', None) else None
    if phase_ms is not None:
        head._timing = None
    barrier()
    # The loop a trainer runs: the H2D copy of step k+1's embeddings goes up on a copy stream while step k computes, and the loss
    # of step k is read back (pinned buffer, async copy) once step k+1 has been enqueued.  Every step's inputs cross PCIe and
    # every step's loss is read on the host inside the timed region.
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]

    def stage(s):
        x, y, xl, yl = pinned[(args.warmup + s) % n_b]
        with torch.cuda.stream(copy_stream):
            xd, yd = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xd, yd, xl, yl, ev

    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    nxt, pending, lv = stage(0), None, 0.0
    for s in range(args.steps):
        xd, yd, xl, yl, ev = nxt
        main.wait_event(ev)
        xd.record_stream(main)
        yd.record_stream(main)
        loss_e = step_api(xd, yd, xl, yl)
        buf = loss_host[s & 1]
        buf.copy_(loss_e.detach().reshape(1), non_blocking=True)
        ev_l = torch.cuda.Event()
        ev_l.record(main)
        if s + 1 < args.steps:
            nxt = stage(s + 1)
            if prefetch:
                prefetch(nxt[2], nxt[3])
        if pending is not None:
            pending[1].synchronize()
            lv = float(pending[0])                    # D2H read of the previous step's loss
        pending = (buf, ev_l)
    pending[1].synchronize()
    lv = float(pending[0])
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    t_all = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_all[0]), float(t_all[1])
    samples = 2 * B * world * args.steps
    value = samples / (ms * 1e-3)
    pk = peaks()
    rows_per_sweep = B * world
    q_local = Q // world
    flops = 4.0 * rows_per_sweep * q_local * D                     # algorithmic FLOPs of one main sweep launch (4*B*Q*D)
    ach = flops * sweep_n / (sweep_ms * 1e-3) / 1e12 if sweep_ms > 0 else 0.0
    if rank == 0:
        cb, _ = cpu_reference(w, 2, 1) if world == 1 and not args.no_cpu else (None, None)
        line = dict(metric='ffc_head_fwd_bwd_samples_per_s', value=value, unit='samples/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16', data='synthetic',
                    config=config_of(w, world),
                    e2e=dict(value=samples / (ms_e2e * 1e-3), unit='samples/s', h2d_bytes_per_step=2 * B * D * 4 + 2 * B * 8, d2h_bytes_per_step=4,
                             ms_per_step=ms_e2e / args.steps, last_loss=lv,
                             pipeline='H2D of step k+1 on a copy stream under step k; loss of step k read on the host after step k+1 is enqueued'),
                    gpu_launches=int(launches),
                    roofline=dict(bound='tensor', achieved=ach, peak=pk['sustained'], unit='TFLOP/s', frac=ach / pk['sustained'], traffic=sweep_traffic(w, world),
                                  kernel='ffc_head_sweep_sm100_kernel (main sweep)', launches=int(sweep_n), avg_ms=sweep_ms / max(1, sweep_n),
                                  algorithmic_flops_per_launch=flops, peak_kind=f'bf16_tflops_sustained ({pk["src"]})',
                                  frac_of_burst=ach / pk['burst'], sweep_share_of_step=sweep_ms / ms),
                    clocks=clk, loss=loss_val)
        if cb is not None:
            line['cpu_baseline'] = cb
        if world > 1 and os.environ.get('FFC_DIST_TIMING'):
            line['phase_ms_total'] = phase_ms
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == '__main__':
    main()
