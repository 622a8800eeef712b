#!/usr/bin/env python
"""Simulation of the hard-negative threshold logic of the D = 512 sweep (csrc/head_sm100.cu) for one warp's 32 rows at the 8-way shard
shape (131 072 local columns N(0, 1/512), k = 10, 8 column-chunk items run one after the other, 3 interleaved warpgroups per item, each
with its own list; the k-th value shared between items is only raised by an item whose own list is full).  Prints, without and with
seeding an item's lists from the finished items' partials: fraction of 32-column chunks with at least one candidate row, candidate rows
per chunk, insertions per chunk, and the per-item trigger rate.  The seeded variant also asserts that the merged result is the exact top-k.
    python tools/topk_threshold_sim.py          (CPU, ~1 min; behind profiles/r2_outlier_sweep.md)"""
import numpy as np

rng=np.random.default_rng(0)
K=10; NCOL=131072; CH=8; TILES=NCOL//128//CH   # tiles per item
ROWS=32
def run_unseeded(seq_items=True):
    trig_chunks=0; tot_chunks=0; cand_lanes=0; hits=0
    floor=np.zeros(ROWS)            # shared threshold per row
    x=rng.standard_normal((ROWS,NCOL)).astype(np.float32)/np.sqrt(512)
    per_item=[]
    for it in range(CH):
        lists=[[[] for _ in range(ROWS)] for _ in range(3)]
        kth=np.zeros((3,ROWS))
        t_it=0; c_it=0
        for t in range(TILES):
            g=t%3
            col0=(it*TILES+t)*128
            fl=floor.copy()          # picked up once per tile
            for cc in range(4):
                blk=x[:,col0+cc*32:col0+cc*32+32]
                thr=np.maximum(kth[g],fl)
                mx=blk.max(axis=1)
                cl=mx>thr
                tot_chunks+=1; c_it+=1
                if cl.any():
                    trig_chunks+=1; t_it+=1
                    cand_lanes+=int(cl.sum())
                    for r in np.nonzero(cl)[0]:
                        for v in blk[r]:
                            if v>max(kth[g][r],fl[r]):
                                L=lists[g][r]; L.append(v); L.sort(reverse=True); del L[K:]
                                hits+=1
                                if len(L)==K: kth[g][r]=L[-1]
            # publish
            for r in range(ROWS):
                if len(lists[g][r])==K: floor[r]=max(floor[r],lists[g][r][-1])
        per_item.append(t_it/c_it)
    return trig_chunks/tot_chunks, cand_lanes/tot_chunks, hits/tot_chunks, per_item



def run_seeded():
    trig=0; tot=0; cand_lanes=0; hits=0
    floor=np.zeros(ROWS)
    x=rng.standard_normal((ROWS,NCOL)).astype(np.float32)/np.sqrt(512)
    partials=[]   # per finished item: list per row of own top-k
    per_item=[]
    for it in range(CH):
        # seed: top-k of union of finished partials
        seed=[]
        for r in range(ROWS):
            u=sorted([v for p in partials for v in p[r]],reverse=True)[:K]
            seed.append(u)
        lists=[[[(v,True) for v in seed[r]] for r in range(ROWS)] for _ in range(3)]
        kth=np.array([[ (lists[g][r][-1][0] if len(lists[g][r])==K else 0.0) for r in range(ROWS)] for g in range(3)])
        t_it=0;c_it=0
        for t in range(TILES):
            g=t%3; col0=(it*TILES+t)*128; fl=floor.copy()
            for cc in range(4):
                blk=x[:,col0+cc*32:col0+cc*32+32]
                thr=np.maximum(kth[g],fl); mx=blk.max(axis=1); cl=mx>thr
                tot+=1;c_it+=1
                if cl.any():
                    trig+=1;t_it+=1; cand_lanes+=int(cl.sum())
                    for r in np.nonzero(cl)[0]:
                        for v in blk[r]:
                            if v>max(kth[g][r],fl[r]):
                                L=lists[g][r]; L.append((v,False)); L.sort(reverse=True); del L[K:]; hits+=1
                                if len(L)==K: kth[g][r]=L[-1][0]
            for r in range(ROWS):
                if len(lists[g][r])==K: floor[r]=max(floor[r],lists[g][r][-1][0])
        own=[sorted([v for g in range(3) for (v,s) in lists[g][r] if not s],reverse=True)[:K] for r in range(ROWS)]
        partials.append(own)
        per_item.append(round(t_it/c_it,3))
    # check exactness
    for r in range(ROWS):
        got=sorted([v for p in partials for v in p[r]],reverse=True)[:K]
        ref=sorted(x[r].tolist(),reverse=True)[:K]
        assert np.allclose(got,ref),(r,got,ref)
    return trig/tot, cand_lanes/tot, hits/tot, per_item



if __name__ == '__main__':
    print('unseeded:', run_unseeded())
    print('seeded:  ', run_seeded())
