"""Training-step harness around the B200 head: the reference's inner loop (main.py:23-86) with the loaders
abstracted, so that it runs on synthetic sources of the reference's batch shapes as well as on the reference's own LMDB loaders.

Nothing here computes: batch composition is tensor concatenation, the step is the reference's AMP sequence around ``ffc_net(x, y, x_label,
y_label)``.  What it adds over main.py is ordering: the labels of a step are split off and handed to the head (`prefetch_labels`, when the
head has one -- the sharded head, ffc_b200/dist.py) before the images go through the backbones, so the label exchange and the LRU
bookkeeping run under the previous step's sweeps and the backbones' kernels.

Host logic: tests/test_train_host.py (CPU, stand-in module).  On a GPU: tests/test_gpu_train.py (a toy backbone ending in FFCTail, and
the reference's own MobileFaceNet / iresnet50 from the staged reference tree, C1 / C2 shapes).
"""
from __future__ import annotations

import os
import time

import torch


def compose_batch(images1, images2, id_indexes, ins_images, instance_label):
    """main.py:49-60: an id batch (two views of B/2 identities) and an instance batch (B images) become the two views of one FFC batch --
    x = [view 1 of the ids | first half of the instances], y = [view 2 | second half], labels alike.  Labels stay where they are (CPU in
    the reference, main.py:59-60); images are moved by the caller."""
    ins1, ins2 = torch.chunk(ins_images, 2)
    lab1, lab2 = torch.chunk(instance_label, 2)
    x = torch.cat([images1, ins1])
    y = torch.cat([images2, ins2])
    x_label = torch.cat([id_indexes, lab1])
    y_label = torch.cat([id_indexes, lab2])
    return x, y, x_label, y_label


class SyntheticSource:
    """Stand-in for the reference's two loaders (main.py:100-109) with their item formats (util/lmdb_loader.py:101-132, 191-237):
    ``instance_loader()`` yields (images [B,3,S,S], labels [B], -1) and ``id_loader()`` yields (view1, view2, id_index) batches of B/2;
    identities are dense class indices in [0, num_class).  Seeded, CPU tensors (pinned when CUDA is available)."""

    def __init__(self, num_class, batch_size, image_size=112, batches_per_epoch=100, seed=0, channels=3):
        assert batch_size % 2 == 0, 'main.py:49-50 splits the instance batch in two'
        self.num_class, self.B, self.S, self.C, self.n, self.seed = num_class, batch_size, image_size, channels, batches_per_epoch, seed
        self._pin = torch.cuda.is_available()

    def _images(self, n, gen):
        t = torch.randn(n, self.C, self.S, self.S, generator=gen)
        return t.pin_memory() if self._pin else t

    def instance_loader(self, epoch=0):
        gen = torch.Generator().manual_seed(self.seed + 7919 * epoch)
        for _ in range(self.n):
            yield self._images(self.B, gen), torch.randint(0, self.num_class, (self.B,), generator=gen), -1

    def id_loader(self, epoch=0):
        gen = torch.Generator().manual_seed(self.seed + 7919 * epoch + 1)
        h = self.B // 2
        while True:
            perm = torch.randperm(self.num_class, generator=gen)        # RandomSampler over identities: distinct within a batch
            for a in range(0, self.num_class - h + 1, h):
                yield self._images(h, gen), self._images(h, gen), perm[a:a + h].clone()

    def __len__(self):
        return self.n


def train_step(ffc_net, optimizer, scaler, x, y, x_label, y_label, autocast_dtype=None, device_type='cuda'):
    """main.py:53-71: zero_grad, forward under autocast, scaled backward, optimiser step, scaler update.  Returns the loss tensor (not
    synchronised: main.py reads it only every 1000 iterations).  ``autocast_dtype=None`` is main.py:64's ``torch.amp.autocast('cuda')``: the
    device's default autocast type (fp16 on CUDA, hence the GradScaler)."""
    optimizer.zero_grad()
    with torch.amp.autocast(device_type, dtype=autocast_dtype):
        loss = ffc_net(x, y, x_label, y_label)
    scaler.scale(loss).backward()
    scaler.step(optimizer)
    scaler.update()
    return loss


def train_one_epoch(id_loader, instance_loader, ffc_net, optimizer, scaler, cur_epoch=1, saved_dir=None, real_iter=0, save_every=1000,
                    lr_scheduler=None, db_size=None, device=None, autocast_dtype=None, log=None, rank=None):
    """main.py:23-86.  ``instance_loader`` drives the epoch; ``id_loader`` is restarted when it runs out (main.py:43-47).  Every
    ``save_every`` iterations the reference's snapshot dict is written (main.py:84-85: probe weights, LRU, queue, queue positions).
    One batch is composed ahead of the step it feeds, so that its labels can be handed to the head early (``prefetch_labels``: both
    ``ffc_b200.FFC`` and ``ShardedFFCHead`` have it).  ``rank``: with a sharded head every rank saves its own shard, as
    ``<n>.rank<r>.pt`` (ShardedFFCHead.checkpoint is per rank)."""
    device = torch.device('cuda') if device is None else torch.device(device)
    id_iter = iter(id_loader)
    prefetch = getattr(ffc_net, 'prefetch_labels', None)
    start = time.time()

    def next_batch(it):
        nonlocal id_iter
        item = next(it, None)
        if item is None:
            return None
        ins_images, instance_label, _ = item
        try:
            images1, images2, id_indexes = next(id_iter)
        except StopIteration:
            id_iter = iter(id_loader)
            images1, images2, id_indexes = next(id_iter)
        x, y, xl, yl = compose_batch(images1, images2, id_indexes, ins_images, instance_label)
        return x.to(device, non_blocking=True), y.to(device, non_blocking=True), xl, yl

    it = iter(instance_loader)
    cur = next_batch(it)
    batch_idx = 0
    loss = None
    while cur is not None:
        if lr_scheduler is not None and db_size:
            lr_scheduler.update(None, batch_idx * 1.0 / db_size)                                   # main.py:40-41
        x, y, xl, yl = cur
        loss = train_step(ffc_net, optimizer, scaler, x, y, xl, yl, autocast_dtype, device.type)
        cur = next_batch(it)                                                                      # composed while the step runs on the device
        if cur is not None and prefetch is not None:
            prefetch(cur[2], cur[3])
        real_iter += 1
        batch_idx += 1
        if save_every and real_iter % save_every == 0:
            if log is not None:
                log(dict(epoch=cur_epoch, iter=real_iter, loss=float(loss), seconds=time.time() - start))
            start = time.time()
            if saved_dir is not None:
                os.makedirs(saved_dir, exist_ok=True)
                name = '%d.pt' % (real_iter // save_every) if rank is None else '%d.rank%d.pt' % (real_iter // save_every, rank)
                torch.save(ffc_net.checkpoint(), os.path.join(saved_dir, name))   # main.py:84-85
    return real_iter
