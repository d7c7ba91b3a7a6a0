"""Model check of the barrier protocol of the fused chain kernels (csrc/sdf_chain.cuh), on the CPU.

The kernel's five roles -- MMA issuer, auxiliary-block loader, storer and the epilogue warps (the weight loader's ring is a
plain two-way producer/consumer pair and is left out) -- talk through parity-waited mbarriers.  A parity wait cannot tell
"phase k has not completed" from "phases k and k+1 have both completed", so the protocol is only correct if no producer can
get two phases ahead of a waiter on any barrier.  Round 2 found one place where it could (a per-slot "block done" barrier:
the epilogue of a step that needs no auxiliary slot ran four blocks ahead of a storer delayed by a cold first launch), as an
intermittent hang on the GPU.  This file restates the protocol as generators over a tiny mbarrier model and runs it under
ADVERSARIAL schedules -- each role in turn only runs when nobody else can -- and random ones:

* the protocol as shipped never deadlocks;
* the pre-fix protocol (block-done barrier per slot, storer not consuming the loader's phase) deadlocks under the
  starved-storer schedule, i.e. the model is sharp enough to see the bug it guards against.

The model mirrors sdf_chain.cuh by hand (role loops: "weight loader / MMA issuer / auxiliary-block loader / storer /
operand builders / epilogue"); change both together."""
import random

import pytest

NSLOT = 2          # sc_nslot<FAM_SDF_FWD / FAM_RELU>() (the backward family has 3: covered by the parametrisation below)
E_WARPS = 4        # epilogue warps in the model (16 in the kernel: only the arrival count changes)
SC_NAR = 6


class Bar:
    """mbarrier: `count` arrivals complete a phase; wait(parity) passes when the phase of that parity has completed."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def passed(self, parity):
        return (self.phase & 1) != (parity & 1)


class Chain:
    """steps: list of dicts(nb=blocks, KB=reduction blocks, chain=bool (operand from the previous step), slot=bool (the
    epilogue reads auxiliary blocks), feeds=bool (next step reads this step's output))."""

    def __init__(self, steps, tiles, nslot, fixed, storer_waits_loader=None):
        self.steps, self.tiles, self.nslot, self.fixed = steps, tiles, nslot, fixed
        self.storer_waits_loader = fixed if storer_waits_loader is None else storer_waits_loader
        nd = 4 if fixed else nslot
        self.a_ready = [Bar(E_WARPS) for _ in range(SC_NAR)]
        self.acc_full, self.acc_half0, self.op_free = Bar(1), Bar(1), Bar(1)
        self.aux_full = [Bar(1) for _ in range(nslot)]
        self.aux_empty = [Bar(1) for _ in range(nslot)]
        self.blk_done = [Bar(E_WARPS) for _ in range(nd)]

    # every role is a generator that yields (barrier, parity) when it has to wait and None after a unit of work
    def mma(self):
        lg = 0
        for _ in range(self.tiles):
            for S in self.steps:
                for kb in range(S["KB"]):
                    yield (self.a_ready[kb], lg & 1)
                    if kb == S["KB"] - 1:
                        self.acc_half0.arrive()              # commit after the first half-tile of the last K-block
                    yield None
                for kb in range(S["KB"], SC_NAR):
                    yield (self.a_ready[kb], lg & 1)
                self.acc_full.arrive()
                lg += 1

    def aux_loader(self):
        c = 0
        for _ in range(self.tiles):
            for S in self.steps:
                for _b in range(S["nb"]):
                    slot = c % self.nslot
                    if c >= self.nslot:
                        yield (self.aux_empty[slot], ((c // self.nslot) - 1) & 1)
                    self.aux_full[slot].arrive()             # the bulk copies' complete_tx, or a plain arrive
                    c += 1
                    yield None

    def storer(self):
        c = 0
        for _ in range(self.tiles):
            for S in self.steps:
                for _b in range(S["nb"]):
                    slot = c % self.nslot
                    if self.storer_waits_loader:
                        yield (self.aux_full[slot], (c // self.nslot) & 1)
                    if self.fixed:
                        yield (self.blk_done[c & 3], (c >> 2) & 1)
                    else:
                        yield (self.blk_done[slot], (c // self.nslot) & 1)
                    self.aux_empty[slot].arrive()
                    c += 1
                    yield None
                self.op_free.arrive()

    def epilogue(self):
        lg = c = nfree = 0
        for _ in range(self.tiles):
            for si, S in enumerate(self.steps):
                first = lg == 0
                if not S["chain"]:
                    if not first:
                        yield (self.op_free, nfree & 1)
                        nfree += 1
                    yield None                               # build the operand
                    for b in self.a_ready:
                        b.arrive()
                yield (self.acc_half0, lg & 1)
                full_done = False
                if S["chain"] and not first:
                    yield (self.op_free, nfree & 1)
                    nfree += 1
                pub = S["feeds"]
                for b in range(S["nb"]):
                    if (b >= 2 or b == S["KB"] - 1) and not full_done:
                        yield (self.acc_full, lg & 1)
                        full_done = True
                    slot = c % self.nslot
                    if S["slot"]:
                        yield (self.aux_full[slot], (c // self.nslot) & 1)
                    yield None                               # the block's arithmetic
                    (self.blk_done[c & 3] if self.fixed else self.blk_done[slot]).arrive()
                    if pub:
                        self.a_ready[b].arrive()
                    c += 1
                if not full_done:
                    yield (self.acc_full, lg & 1)
                if pub:
                    for b in range(S["nb"], SC_NAR):
                        self.a_ready[b].arrive()
                lg += 1


def run(chain, policy, rng=None, max_ticks=200000):
    """policy: ('starve', role_index) -> that role only runs when no other role can; ('random',) -> uniform choice.
    Returns True when every role finished, False on deadlock."""
    roles = [chain.mma(), chain.aux_loader(), chain.storer()] + [chain.epilogue() for _ in range(E_WARPS)]
    waiting = [None] * len(roles)
    done = [False] * len(roles)

    def runnable(i):
        if done[i]:
            return False
        w = waiting[i]
        return w is None or w[0].passed(w[1])

    for _ in range(max_ticks):
        if all(done):
            return True
        cand = [i for i in range(len(roles)) if runnable(i)]
        if not cand:
            return False
        if policy[0] == "starve":
            others = [i for i in cand if i != policy[1]]
            i = (others[0] if rng is None else rng.choice(others)) if others else policy[1]
        else:
            i = rng.choice(cand)
        waiting[i] = None
        try:
            waiting[i] = next(roles[i])
        except StopIteration:
            done[i] = True
    raise AssertionError("model did not terminate")


def _relu_forward(n_layers=5):
    """The colour network's forward chain: ReLU layers that need no slot, then a narrow output step that does."""
    steps = [dict(nb=4, KB=5 if l == 0 else 4, chain=l > 0, slot=False, feeds=True) for l in range(n_layers - 1)]
    return steps + [dict(nb=1, KB=4, chain=True, slot=True, feeds=False)]


def _sdf_forward():
    """Value chain (no slot) + feature step + tangent chain (h blocks through the slots) + normal step."""
    steps = [dict(nb=4, KB=1 if l == 0 else 4, chain=l > 0, slot=False, feeds=True) for l in range(8)]
    steps.append(dict(nb=4, KB=4, chain=True, slot=True, feeds=True))                    # FEATQ
    steps += [dict(nb=4, KB=4, chain=True, slot=True, feeds=True) for _ in range(7)]     # SPMUL
    steps.append(dict(nb=1, KB=4, chain=True, slot=True, feeds=False))                   # G0
    return steps


def _nerf_forward():
    """A narrow output step in the MIDDLE of a chain (alpha), then more layers."""
    steps = [dict(nb=4, KB=2 if l == 0 else 4, chain=l > 0, slot=False, feeds=True) for l in range(8)]
    steps.append(dict(nb=1, KB=4, chain=True, slot=True, feeds=True))                    # alpha: leaves the operand in place
    steps += [dict(nb=4, KB=4, chain=True, slot=False, feeds=True) for _ in range(2)]
    steps.append(dict(nb=1, KB=2, chain=True, slot=True, feeds=False))
    return steps


CHAINS = {"relu_forward": _relu_forward, "sdf_forward": _sdf_forward, "nerf_forward": _nerf_forward}


@pytest.mark.parametrize("nslot", [2, 3])
@pytest.mark.parametrize("name", sorted(CHAINS))
def test_shipped_protocol_never_deadlocks(name, nslot):
    steps = CHAINS[name]()
    n_roles = 3 + E_WARPS
    for starved in range(n_roles):                      # each role in turn runs only when nobody else can
        assert run(Chain(steps, 3, nslot, fixed=True), ("starve", starved)), "deadlock with role %d starved" % starved
    rng = random.Random(7)
    for trial in range(60):
        pol = ("starve", rng.randrange(n_roles)) if trial % 2 else ("random",)
        assert run(Chain(steps, 3, nslot, fixed=True), pol, rng), "deadlock in random trial %d" % trial


def test_model_sees_the_round2_bug():
    """Block-done barrier per slot (two slots) + a storer that lags: the epilogue of a slot-free step completes two phases of
    blk_done[slot] before the storer looks -- the storer's parity wait never returns, and the epilogue waits for op_free."""
    steps = _relu_forward()
    assert not run(Chain(steps, 2, 2, fixed=False), ("starve", 2)), "the pre-fix protocol should deadlock with a starved storer"
    # the same protocol survives when the storer keeps up (why the bug hid for a whole round)
    assert run(Chain(steps, 2, 2, fixed=False), ("starve", 0))


def test_both_halves_of_the_fix_are_needed():
    """Four block-done barriers alone leave the mirror-image hazard: with slot-free steps nothing ties the storer to the
    auxiliary loader, so a lagging LOADER is lapped on aux_empty[slot] by a storer that releases the slot twice.  The storer
    therefore consumes the loader's phase of the slot before it releases it."""
    for name in sorted(CHAINS):
        steps = CHAINS[name]()
        assert not run(Chain(steps, 3, 2, fixed=True, storer_waits_loader=False), ("starve", 1)), name
        assert run(Chain(steps, 3, 2, fixed=True, storer_waits_loader=True), ("starve", 1)), name
