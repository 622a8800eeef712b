"""A model of the synchronisation protocol of the CTA-pair sweep kernel (csrc/head_sm100.cu) -- TEST INFRASTRUCTURE, no GPU.

The kernel's correctness rests on ~60 mbarriers and 4 named barriers shared by nine warp roles of two CTAs, with phases that run on
across work items (persistent launches), items without column tiles and hard-negative-only items that the O-CTA sits out.  A wrong
parity or a barrier that can get two phases ahead of a waiter does not show up as a wrong number but as a hang on a GPU box.  This
file restates the protocol -- every wait / arrive / commit / named barrier of every role, with the kernel's own phase arithmetic --
as coroutines over mbarrier objects with the hardware's semantics (`try_wait.parity p` succeeds iff the barrier's current phase
parity differs from p, so a waiter that is lapped by two phases blocks for ever), runs them under random interleavings with
asynchronous completions (TMA, tcgen05.commit in issue order, st.async) delivered at random later times, and checks

  * progress: every role finishes (no deadlock, no lapped waiter), for any list of items;
  * the resources the barriers protect: a W ring slot, an S accumulator, a P~ buffer, the probe tile P and the O accumulator are
    only written when free and only read when complete, and the O write-out's staging buffer (the P~ buffer of the item's last
    tile) is not touched by the S-CTA until the write-out warps have released it.

It mirrors the kernel source role by role (same variable names); when the kernel's protocol changes, this model has to change
with it -- tests/test_sweep_protocol_model.py runs it, and deliberately broken variants, on the CPU.
"""
import random

NS1, NS2, NSB, NPB, NEPI, NKC, NJB = 12, 4, 2, 3, 3, 8, 4      # D = 512: 8 K chunks per tile, 4 O-CTA stages per tile


class Deadlock(Exception):
    pass


class Hazard(Exception):
    pass


class MBar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f'{self.name}: more arrivals than the barrier counts'
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def passed(self, parity):
        return (self.phase & 1) != parity


class Sim:
    """variant: None = the kernel's protocol; otherwise the name of a deliberate defect (see the test)."""

    def __init__(self, items, seed, variant=None):
        self.items = [it for it in items]                     # [(n_tiles, all_out)]
        self.rng = random.Random(seed)
        self.variant = variant
        b = lambda n, c=1: MBar(n, c)
        self.p_full = b('p_full', NEPI)                       # (the kernel counts 12 warps; a warpgroup is one agent here)
        self.w_full = [b(f'w_full{i}') for i in range(NS1)]
        self.w_empty = [b(f'w_empty{i}') for i in range(NS1)]
        self.s_full = [b(f's_full{i}') for i in range(NEPI)]
        self.s_empty = [b(f's_empty{i}') for i in range(NSB)]
        self.pt_empty = [b(f'pt_empty{i}') for i in range(NPB)]
        self.w2_full = [b(f'w2_full{i}') for i in range(NS2)]
        self.w2_empty = [b(f'w2_empty{i}') for i in range(NS2)]
        self.pt_full = [b(f'pt_full{i}', 2) for i in range(NPB)]      # the hand-off warp's arrive.expect_tx + the tile's bytes
        self.pt_ready = [b(f'pt_ready{i}') for i in range(NPB)]
        self.o_full, self.o_empty = b('o_full'), b('o_empty')
        # resources
        self.w1 = ['free'] * NS1
        self.w2 = ['free'] * NS2
        self.S = ['free'] * NSB
        self.Pt = ['free'] * NPB
        self.P = 'free'            # free -> loaded (all three warpgroups stored their pieces) -> in use
        self.P_pieces = 0
        self.gemm1_open = 0        # GEMM-1 tiles of the current item issued but not yet complete
        self.O = 'drained'
        self.named = {}            # 'name#use' -> [arrived agents]
        self.sync_use = {}         # (agent, name) -> uses so far
        self.pending_async = []    # unordered completions (TMA, st.async)
        self.fifo = {}             # per issuing thread: tcgen05.commit completions, in order

    # -- helpers the roles use -------------------------------------------------------------------------------------------------
    def need(self, cond, msg):
        if not cond:
            raise Hazard(msg)

    def later(self, fn):
        self.pending_async.append(fn)

    def commit(self, who, fn):
        self.fifo.setdefault(who, []).append(fn)

    # -- S-CTA -----------------------------------------------------------------------------------------------------------------
    def s_producer(self):
        stage, ph = 0, 0
        for n_tiles, _ in self.items:
            if n_tiles == 0:
                continue
            for _t in range(n_tiles):
                for _kc in range(NKC):
                    yield ('wait', self.w_empty[stage], ph ^ 1)
                    self.need(self.w1[stage] == 'free', f'W ring slot {stage} overwritten before GEMM-1 has read it')
                    self.w1[stage] = 'loading'

                    def done(stage=stage):
                        self.w1[stage] = 'full'
                        self.w_full[stage].arrive()
                    self.later(done)
                    stage += 1
                    if stage == NS1:
                        stage, ph = 0, ph ^ 1

    def s_mma(self):
        stage, ph, gi, wg, ni = 0, 0, 0, 0, 0
        for n_tiles, _ in self.items:
            if n_tiles == 0:
                continue
            yield ('wait', self.p_full, ni & 1)
            self.need(self.P == 'loaded', 'GEMM-1 issued before the probe tile is complete in TMEM')
            for _i in range(n_tiles):
                sb = gi & (NSB - 1)
                yield ('wait', self.s_empty[sb], ((gi // NSB) & 1) ^ 1)
                self.need(self.S[sb] == 'free', f'S accumulator {sb} overwritten before the epilogue has read it')
                self.S[sb] = 'accumulating'
                self.gemm1_open += 1
                for kc in range(NKC):
                    yield ('wait', self.w_full[stage], ph)
                    self.need(self.w1[stage] == 'full', f'GEMM-1 reads W ring slot {stage} before its TMA has landed')

                    def freed(stage=stage):
                        self.w1[stage] = 'free'
                        self.w_empty[stage].arrive()
                    self.commit('s_mma', freed)
                    if kc == NKC - 1:
                        def full(sb=sb, wg=wg):
                            self.S[sb] = 'full'
                            self.gemm1_open -= 1
                            self.s_full[wg].arrive()
                        self.commit('s_mma', full)
                    stage += 1
                    if stage == NS1:
                        stage, ph = 0, ph ^ 1
                gi += 1
                wg = 0 if wg + 1 == NEPI else wg + 1
            ni += 1

    def s_epilogue(self, g):
        s_use, gi0, hi0 = 0, 0, 0
        for n_tiles, all_out in self.items:
            if n_tiles == 0:
                continue
            # probe tile -> TMEM (this warpgroup's pieces)
            self.need(self.P != 'loaded' or self.variant == 'p_reload_without_barrier', 'probe tile stored while the previous one is still the A operand')
            self.need(self.gemm1_open == 0, 'probe tile overwritten while a GEMM-1 that reads it is in flight')
            self.P_pieces += 1
            if self.P_pieces == NEPI:
                self.P, self.P_pieces = 'loaded', 0
            yield ('arrive', self.p_full)
            first = (g + NEPI - gi0 % NEPI) % NEPI
            for i in range(first, n_tiles, NEPI):
                sb = (gi0 + i) & (NSB - 1)
                yield ('wait', self.s_full[g], s_use & 1)
                hi = hi0 + i
                pb = hi % NPB
                hand = not all_out
                if hand:
                    yield ('wait', self.pt_empty[pb], ((hi // NPB) & 1) ^ 1)
                self.need(self.S[sb] == 'full', f'epilogue reads S accumulator {sb} before its GEMM-1 has completed')
                yield ('step',)                               # the tcgen05.ld of the tile's chunks
                self.S[sb] = 'free'
                yield ('arrive', self.s_empty[sb])
                if hand:
                    self.need(self.Pt[pb] == 'free', f'P~ buffer {pb} written while the O-CTA still owns it (GEMM-2 operand or write-out staging)')
                    self.Pt[pb] = 'writing'

                    def landed(pb=pb):
                        self.Pt[pb] = 'full'
                        self.pt_full[pb].arrive()
                    self.later(landed)
                s_use += 1
            gi0 += n_tiles
            if not all_out:
                hi0 += n_tiles
            if self.variant != 'p_reload_without_barrier':
                yield ('sync', 'bar1a', NEPI)                 # every warpgroup has seen its last s_full: every GEMM-1 of the item is complete
                if g == 0:
                    self.need(self.gemm1_open == 0, 'end-of-item barrier passed with a GEMM-1 of the item outstanding')
                    self.P = 'free'
                yield ('sync', 'bar1b', NEPI)
            else:
                self.P = 'free'

    # -- O-CTA -----------------------------------------------------------------------------------------------------------------
    def o_producer(self):
        stage, ph = 0, 0
        for n_tiles, all_out in self.items:
            if n_tiles == 0 or all_out:
                continue
            for _t in range(n_tiles):
                for _jb in range(NJB):
                    yield ('wait', self.w2_empty[stage], ph ^ 1)
                    self.need(self.w2[stage] == 'free', f'O-CTA W stage {stage} overwritten before GEMM-2 has read it')
                    self.w2[stage] = 'loading'

                    def done(stage=stage):
                        self.w2[stage] = 'full'
                        self.w2_full[stage].arrive()
                    self.later(done)
                    stage += 1
                    if stage == NS2:
                        stage, ph = 0, ph ^ 1

    def o_mma(self):
        stage, ph, pb, pt_ph, ni = 0, 0, 0, 0, 0
        for n_tiles, all_out in self.items:
            if n_tiles == 0 or all_out:
                continue
            if ni > 0 and self.variant != 'no_o_empty':
                yield ('wait', self.o_empty, (ni - 1) & 1)
            self.need(self.O == 'drained', 'GEMM-2 of the next item overwrites O before the write-out warps have read it')
            self.O = 'accumulating'
            for i in range(n_tiles):
                yield ('wait', self.pt_ready[pb], pt_ph)
                self.need(self.Pt[pb] == 'ready', f'GEMM-2 reads P~ buffer {pb} before it is complete and fenced')
                release_pt = i + 1 < n_tiles or self.variant == 'release_last_tile_by_commit'
                for jb in range(NJB):
                    yield ('wait', self.w2_full[stage], ph)
                    self.need(self.w2[stage] == 'full', f'GEMM-2 reads W stage {stage} before its TMA has landed')

                    def freed(stage=stage):
                        self.w2[stage] = 'free'
                        self.w2_empty[stage].arrive()
                    self.commit('o_mma', freed)
                    if jb == NJB - 1:
                        def consumed(pb=pb, release=release_pt):
                            self.Pt[pb] = 'free' if release else 'consumed'
                            if release:
                                self.pt_empty[pb].arrive()
                        self.commit('o_mma', consumed)
                    stage += 1
                    if stage == NS2:
                        stage, ph = 0, ph ^ 1
                pb += 1
                if pb == NPB:
                    pb, pt_ph = 0, pt_ph ^ 1

            def o_done():
                self.O = 'complete'
                self.o_full.arrive()
            self.commit('o_mma', o_done)
            yield ('wait', self.o_full, ni & 1)
            yield ('sync', 'bar2', 2)
            ni += 1

    def o_handoff(self):
        pb, pt_ph = 0, 0
        for n_tiles, all_out in self.items:
            if n_tiles == 0 or all_out:
                continue
            for _i in range(n_tiles):
                yield ('arrive', self.pt_full[pb])            # arrive.expect_tx
                yield ('wait', self.pt_full[pb], pt_ph)
                self.need(self.Pt[pb] == 'full', f'hand-off passes P~ buffer {pb} on before its bytes have landed')
                self.Pt[pb] = 'ready'                         # fence.proxy.async
                yield ('arrive', self.pt_ready[pb])
                pb += 1
                if pb == NPB:
                    pb, pt_ph = 0, pt_ph ^ 1

    def o_writeout(self):
        gtiles = 0
        for n_tiles, all_out in self.items:
            if n_tiles == 0 or all_out:
                continue
            yield ('sync', 'bar2', 2)
            self.need(self.O == 'complete', 'O read out before every tcgen05.mma of the item has completed')
            pb_last = (gtiles + n_tiles - 1) % NPB
            if self.variant != 'release_last_tile_by_commit':
                self.need(self.Pt[pb_last] == 'consumed', f'write-out stages O through P~ buffer {pb_last}, which is {self.Pt[pb_last]}')
            for _c in range(48):                              # the transposes through the staging buffer (a write-out lasts 2-3 tile times)
                yield ('step',)
                self.need(self.Pt[pb_last] in ('consumed',) or self.variant == 'release_last_tile_by_commit' and self.Pt[pb_last] == 'free',
                          f'the S-CTA wrote P~ buffer {pb_last} while the O write-out was staging through it')
            self.O = 'drained'
            yield ('arrive', self.o_empty)
            if self.variant != 'release_last_tile_by_commit':
                self.Pt[pb_last] = 'free'
                yield ('arrive', self.pt_empty[pb_last])      # remote arrive into the S-CTA
            gtiles += n_tiles

    # -- scheduler -------------------------------------------------------------------------------------------------------------
    def run(self, max_events=2_000_000):
        agents = {'s_producer': self.s_producer(), 's_mma': self.s_mma(), 'o_producer': self.o_producer(), 'o_mma': self.o_mma(),
                  'o_handoff': self.o_handoff(), 'o_writeout': self.o_writeout()}
        for g in range(NEPI):
            agents[f's_epilogue{g}'] = self.s_epilogue(g)
        blocked = {}                                          # agent -> pending request
        for name, gen in list(agents.items()):
            blocked[name] = self._advance(name, gen)
        agents = {n: g for n, g in agents.items() if blocked[n] is not None}
        events = 0
        while agents:
            events += 1
            if events > max_events:
                raise Deadlock('event budget exceeded')
            choices = []
            for name in agents:
                req = blocked[name]
                if req[0] == 'wait' and req[1].passed(req[2]):
                    choices.append(('agent', name))
                elif req[0] in ('arrive', 'step'):
                    choices.append(('agent', name))
                elif req[0] == 'sync' and len(self.named.get(req[1], [])) == req[2]:
                    choices.append(('agent', name))
            for i in range(len(self.pending_async)):
                choices.append(('async', i))
            for who, q in self.fifo.items():
                if q:
                    choices.append(('commit', who))
            if not choices:
                raise Deadlock('no role can make progress: ' + ', '.join(f'{n} at {self._show(blocked[n])}' for n in agents))
            kind, key = self.rng.choice(choices)
            if kind == 'async':
                self.pending_async.pop(key)()
            elif kind == 'commit':
                self.fifo[key].pop(0)()
            else:
                name, req = key, blocked[key]
                if req[0] == 'arrive':
                    req[1].arrive()
                nxt = self._advance(name, agents[name])
                if nxt is None:
                    del agents[name]
                blocked[name] = nxt
        return events

    def _advance(self, name, gen):
        """Run the agent to its next request; a named-barrier request registers the agent as arrived at its k-th use of that barrier (it
        resumes once all parties of that use have arrived)."""
        try:
            req = next(gen)
        except StopIteration:
            return None
        if req[0] == 'sync':
            k = self.sync_use.get((name, req[1]), 0)
            self.sync_use[(name, req[1])] = k + 1
            key = f'{req[1]}#{k}'
            self.named.setdefault(key, []).append(name)
            return ('sync', key, req[2])
        return req

    @staticmethod
    def _show(req):
        if req[0] == 'wait':
            return f'wait {req[1].name} parity {req[2]} (phase {req[1].phase})'
        return str(req[0:2])


def random_items(rng, n_items, max_tiles=9):
    items = []
    for _ in range(n_items):
        r = rng.random()
        n = 0 if r < 0.12 else rng.randint(1, max_tiles)
        items.append((n, rng.random() < 0.3))
    return items
