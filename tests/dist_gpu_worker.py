"""NCCL worker for tests/test_gpu_dist.py: R ranks, one GPU each, sharded head vs the sharded oracle."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'very-large-scale-face-recognition_b200'), os.path.join(ROOT, 'tests')]


def main():
    from ffc_b200.dist import ShardedFFCHead
    from test_dist_cpu import ShardedOracle
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    for precision, tol, D, Q, B, n_ids, loss_type, margin in (('bf16', 1e-2, 128, 1024, 96, 1500, 'Arc', 0.5),
                                                               ('fp32', 1e-5, 64, 512, 40, 700, 'AM', 0.4),
                                                               ('bf16', 1e-2, 512, 4096, 128, 4096, 'AM', 0.4),
                                                               ('bf16', 1e-2, 128, 1024, 96, 1500, 'SV', 0.4),
                                                               ('fp32', 1e-5, 64, 512, 40, 700, 'SV', 0.4)):
        torch.manual_seed(0)
        q0 = F.normalize(torch.rand(2, Q, D), dim=2)
        head = ShardedFFCHead(D, Q, 32.0, loss_type, margin, precision=precision, max_batch=B, device=dev)
        Ql = Q // world
        head.backend.set_queue(q0[:, rank * Ql:(rank + 1) * Ql])
        oracle = ShardedOracle(world, D, Q, 32.0, loss_type, margin, queue=q0, dtype=torch.float64)
        gen = torch.Generator().manual_seed(5)
        cen = F.normalize(torch.randn(n_ids, D, generator=gen))
        for s in range(6):
            n = world * B
            xl = torch.randint(0, n_ids, (n,), generator=gen)
            yl = torch.cat([xl[:n // 2], torch.randint(0, n_ids, (n - n // 2,), generator=gen)])
            x = F.normalize(cen[xl] + 0.7 * torch.randn(n, D, generator=gen))
            y = F.normalize(cen[yl] + 0.7 * torch.randn(n, D, generator=gen))
            sl = slice(rank * B, (rank + 1) * B)
            xs = x[sl].to(dev).requires_grad_(True)
            ys = y[sl].to(dev).requires_grad_(True)
            if s == 2:      # the overlapped two-pass entry used by bench.py (commit bookkeeping under the rollback sweep)
                loss, gx, gy = head.forward_pair(xs.detach(), ys.detach(), xl[sl], yl[sl])
                xs.grad, ys.grad = gx, gy
            elif s == 3:    # labels handed over early (CPU tensors): the rollback bookkeeping runs on the bookkeeping stream
                head.prefetch(xl[sl], yl[sl])
                assert head._pre is not None and head.prefetch_hits == 0
                loss, gx, gy = head.forward_pair(xs.detach(), ys.detach())
                assert head.prefetch_hits == 1
                xs.grad, ys.grad = gx, gy
            elif s == 4:    # a stale prefetch (other labels) is discarded without a trace, then the same objects are accepted
                mine = (xl[sl], yl[sl])
                head.prefetch(yl[sl].clone(), xl[sl].clone())
                head.prefetch(*mine)
                loss, gx, gy = head.forward_pair(xs.detach(), ys.detach(), *mine)
                assert head._pre is None and head.prefetch_hits == 2
                xs.grad, ys.grad = gx, gy
            elif s == 5:    # stale prefetch followed by the autograd entry
                head.prefetch(yl[sl].clone(), xl[sl].clone())
                loss = head.forward(xs, ys, xl[sl], yl[sl])
                loss.backward()
                assert head._pre is None and head.prefetch_hits == 2
            else:
                loss = head.forward(xs, ys, xl[sl], yl[sl])
                loss.backward()
            xo, yo = x.double().requires_grad_(True), y.double().requires_grad_(True)
            ref = oracle.forward(xo, yo, xl.tolist(), yl.tolist())
            ref.backward()
            assert head._last['label'].tolist() == oracle.trace[-1]['labels'], (precision, s, 'labels')
            assert head.backend.lru.state_dict() == oracle.lrus[rank].state_dict(), (precision, s, 'lru')
            assert abs(float(loss) - float(ref)) <= tol * abs(float(ref)), (precision, s, float(loss), float(ref))
            for got, want in ((xs.grad, xo.grad[sl]), (ys.grad, yo.grad[sl])):
                err = float((got.double().cpu() - want).norm() / want.norm())
                assert err <= tol, (precision, s, err)
        del head
    dist.barrier()
    if rank == 0:
        print('DIST_GPU_OK', world)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
