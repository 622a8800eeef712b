#!/usr/bin/env python
"""Everything written after the round's GPU budget was spent, in one run, for the first GPU call of the next round.

    python tools/gpu_checks_pending.py                                               # one GPU
    torchrun --nproc-per-node 2 tools/gpu_checks_pending.py --dist                   # two GPUs: sharded checkpoint / resume, all-hit probe

One GPU:  (1) the C1-shaped fixtures (tests/golden/c1_*.npz) through the CUDA head, fp32 check mode (1e-5) and bf16 (1e-2), bookkeeping and
final queue bit-exact;  (2) ffc_b200.train on a toy backbone ending in FFCTail: 12 steps, finite losses, snapshots written in the
reference's format and resumed;  (3) bench workload c4 is `python bench.py --workload c4` (not run from here).
Prints one line per check and exits non-zero on the first failure.  Each passing check should then move into tests/ as an `-m gpu` test."""
import argparse
import glob
import hashlib
import os
import sys
import tempfile

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, 'very-large-scale-face-recognition_b200'), os.path.join(R, 'tests')]
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def seeded_queue(Q, D, seed):
    return F.normalize(torch.rand(2, Q, D, generator=torch.Generator().manual_seed(seed)), dim=2)


def check_c1(path, precision, tol):
    import ffc_b200
    z = np.load(path)
    dev = torch.device('cuda')
    D, Q, B = int(z['D']), int(z['Q']), int(z['B'])
    m = ffc_b200.FFC('identity', D, queue_size=Q, scale=float(z['scale']), loss_type=str(z['loss_type']), margin=float(z['margin']),
                     precision=precision, max_batch=B)
    m.queue.copy_(seeded_queue(Q, D, int(z['queue_seed'])))
    m = m.to(dev)
    m.lru.restore([(i, i) for i in range(int(z['warm']))])
    for s in range(int(z['steps'])):
        x = torch.from_numpy(z[f'x{s}']).to(dev).requires_grad_(True)
        y = torch.from_numpy(z[f'y{s}']).to(dev).requires_grad_(True)
        xl, yl = torch.from_numpy(z[f'xl{s}']), torch.from_numpy(z[f'yl{s}'])
        px = F.normalize(x)
        loss2 = m.head(px, F.normalize(y).detach(), xl, yl, commit=False)
        rb = m.last_bookkeeping()
        py = F.normalize(y)
        loss1 = m.head(py, F.normalize(x).detach(), yl, xl, commit=True)
        cm = m.last_bookkeeping()
        gx, gy = torch.autograd.grad(loss1 + loss2, [px, py])
        for got, pn in ((rb, 'rb'), (cm, 'cm')):
            for val, k in zip(got, ('rows', 'cols', 'labels', 'ones')):
                assert val == z[f'{pn}_{k}{s}'].tolist(), (s, pn, k)
        assert [list(kv) for kv in m.lru.state_dict()] == z[f'lru{s}'].tolist(), (s, 'lru')
        assert [m.queue_position_dict[i] for i in range(Q)] == z[f'qpos{s}'].tolist(), (s, 'qpos')
        ref = float(z[f'loss{s}'])
        assert abs(float(loss1 + loss2) - ref) <= max(tol, 2e-5) * abs(ref), (s, float(loss1 + loss2), ref)
        assert rel(gx.cpu(), torch.from_numpy(z[f'dx{s}'])) <= max(tol, 3e-5), (s, 'dx')
        assert rel(gy.cpu(), torch.from_numpy(z[f'dy{s}'])) <= max(tol, 3e-5), (s, 'dy')
    sha = hashlib.sha256(m.queue.detach().cpu().contiguous().numpy().tobytes()).hexdigest()
    # F.normalize on the GPU and on the CPU may differ in the last bit, and the enqueued rows are copies of the GPU's: report, do not assert
    return 'queue sha256 ' + ('matches the CPU reference' if sha == str(z['queue_final_sha256']) else 'differs (GPU vs CPU normalisation rounding)')


def check_train():
    import ffc_b200
    from ffc_b200 import train as T
    dev = torch.device('cuda')
    D, Q, B, S = 64, 512, 32, 8

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(3 * S * S, D)
            self.features = ffc_b200.FFCTail(D)

        def forward(self, x):
            return self.features(self.fc(torch.flatten(x, 1)).float())

    torch.manual_seed(0)
    net = ffc_b200.FFC('x', D, queue_size=Q, loss_type='AM', margin=0.4, probe_net=Net(), gallery_net=Net(), max_batch=B).to(dev)
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.05, momentum=0.9)
    scaler = torch.amp.GradScaler('cuda')
    src = T.SyntheticSource(num_class=300, batch_size=B, image_size=S, batches_per_epoch=12, seed=2)
    logs = []
    with tempfile.TemporaryDirectory() as d:
        n = T.train_one_epoch(src.id_loader(), src.instance_loader(), net, opt, scaler, saved_dir=d, save_every=4, device=dev, log=logs.append)
        assert n == 12 and sorted(os.listdir(d)) == ['1.pt', '2.pt', '3.pt'], os.listdir(d)
        ck = torch.load(os.path.join(d, '3.pt'), weights_only=False)
    assert all(np.isfinite(l['loss']) for l in logs), logs
    assert len(ck['lru']) == net.lru.cur_idx > 0 and tuple(ck['fc'].shape) == (2, Q, D)
    net2 = ffc_b200.FFC('x', D, queue_size=Q, loss_type='AM', margin=0.4, probe_net=Net(), gallery_net=Net(), max_batch=B).to(dev)
    net2.load_checkpoint(ck)
    assert net2.lru.state_dict() == net.lru.state_dict() and torch.equal(net2.queue, net.queue)
    return 'losses ' + ' '.join(f"{l['loss']:.3f}" for l in logs)


def check_dist():
    import io
    import torch.distributed as dist
    from ffc_b200.dist import ShardedFFCHead
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    D, Q, B, N = 128, 4096, 64, 6000
    gen = torch.Generator().manual_seed(40 + rank)

    def batch():
        xl = torch.randint(0, N, (B,), generator=gen)
        yl = torch.cat([xl[:B // 2], torch.randint(0, N, (B - B // 2,), generator=gen)])
        return F.normalize(torch.randn(B, D, generator=gen)).to(dev), F.normalize(torch.randn(B, D, generator=gen)).to(dev), xl, yl
    torch.manual_seed(1)
    a = ShardedFFCHead(D, Q, 32.0, 'Arc', 0.5, max_batch=B, device=dev)
    for _ in range(4):
        a.forward_pair(*batch())
    buf = io.BytesIO()
    torch.save(a.checkpoint(), buf)
    buf.seek(0)
    b = ShardedFFCHead(D, Q, 32.0, 'Arc', 0.5, max_batch=B, device=dev)
    b.load_checkpoint(torch.load(buf, weights_only=False))
    for _ in range(3):
        bt = batch()
        for u, v in zip(a.forward_pair(*bt), b.forward_pair(*bt)):
            assert torch.equal(u, v)
        assert a.backend.lru.state_dict() == b.backend.lru.state_dict() and torch.equal(a.backend.queue, b.backend.queue)
    dist.barrier()
    if rank == 0:
        print('ok   sharded checkpoint / resume on', world, 'GPUs: bit-identical continuation')
    dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--dist', action='store_true')
    args = ap.parse_args()
    if args.dist:
        check_dist()
        sys.exit(0)
    for path in sorted(glob.glob(os.path.join(R, 'tests', 'golden', 'c1_*.npz'))):
        for precision, tol in (('fp32', 1e-5), ('bf16', 1e-2)):
            print('ok  ', os.path.basename(path), precision, check_c1(path, precision, tol))
    print('ok   train loop:', check_train())
