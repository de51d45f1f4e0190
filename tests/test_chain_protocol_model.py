"""A host-side MODEL of the work-item queue of the timestep chains (csrc/emc_flight.cu,
flight_scatter_kernel<..., CHAIN>; ChainCtl in csrc/emc_host.h), run under random interleavings.

The kernel itself is tested on the GPU (tests/test_gpu_parity.py::test_timestep_chain_*).  This model
checks the ARGUMENT the kernel's comments make, for any schedule and any number of resident blocks:

* every (slice, timestep) item is processed exactly once, the timesteps of a slice in order, and a
  timestep of a slice starts only after the previous one of THAT slice has been written back;
* the pushes (G implicit initial items + one per finished non-final item) match the K*G consumed
  tickets one to one, every block exits, nothing waits for ever -- also when only some of the G
  blocks are resident at a time (a ticket is only ever held by a running block);
* the ring: a reader empties its slot, a pusher waits for an empty slot, a reader accepts only its own tag.
  Without the handshake a ring of any size is only safe "in practice": a reader (or a pusher) that has taken its
  ticket and is then held up while R more items are pushed would find its slot overwritten (or overwrite a newer
  item) -- this model found exactly that schedule for R = G + 1 in its first version.  With it, G + 1 slots are
  enough and nothing is ever lost, whatever the schedule.
"""
import random

import pytest


def run_chain(G, K, resident, ring, seed):
    rng = random.Random(seed)
    total = K * G
    popped = pushed = 0
    ready = [0] * ring                       # (tag << 32) | (step << 16) | slice
    written = [0] * G                        # timesteps of a slice written back so far
    log = []
    # block states: ("idle",) not yet started, ("pop",), ("wait", T), ("work", slice, j, left), ("push", slice, j),
    # ("pushwait", slice, j, P), ("exit",)
    blocks = [("idle",)] * G
    started = 0
    steps = 0
    while any(b[0] != "exit" for b in blocks):
        steps += 1
        assert steps < 200 * (total + G) + 1000, "no progress: deadlock"
        running = [i for i, b in enumerate(blocks) if b[0] not in ("idle", "exit")]
        # the hardware starts the next block of the grid whenever a resident slot is free
        if len(running) < resident and started < G:
            blocks[started] = ("pop",)
            started += 1
            continue
        i = rng.choice(running)
        b = blocks[i]
        if b[0] == "pop":
            T = popped
            popped += 1
            if T >= total:
                blocks[i] = ("exit",)
            elif T < G:
                blocks[i] = ("work", T, 0, rng.randint(1, 4))
            else:
                blocks[i] = ("wait", T)
        elif b[0] == "wait":
            T = b[1]
            v = ready[T % ring]
            if (v >> 32) == T + 1:
                ready[T % ring] = 0                   # the reader empties the slot
                blocks[i] = ("work", v & 0xffff, (v >> 16) & 0xffff, rng.randint(1, 4))
            # (any other tag -- an empty slot, or a LATER item whose pusher overtook this one's: it waits on)
        elif b[0] == "work":
            _, s, j, left = b
            if left == rng.randint(1, 4) or left == 1:
                assert written[s] == j, "timestep %d of slice %d started before timestep %d was written back" % (j, s, j - 1)
            blocks[i] = ("work", s, j, left - 1) if left > 1 else ("push", s, j)
        elif b[0] == "push":
            _, s, j = b
            assert written[s] == j
            written[s] = j + 1
            log.append((s, j))
            if j + 1 < K:
                P = G + pushed
                pushed += 1
                blocks[i] = ("pushwait", s, j, P)
            else:
                blocks[i] = ("pop",)
        elif b[0] == "pushwait":
            _, s, j, P = b
            if ready[P % ring] == 0:                  # the pusher waits for an empty slot
                ready[P % ring] = ((P + 1) << 32) | ((j + 1) << 16) | s
                blocks[i] = ("pop",)
    assert popped == total + G and pushed == total - G
    assert sorted(log) == [(s, j) for s in range(G) for j in range(K)]
    assert written == [K] * G
    return steps


@pytest.mark.parametrize("G,K,resident", [(1, 1, 1), (1, 9, 1), (7, 5, 7), (7, 5, 3), (7, 5, 1), (16, 12, 16),
                                            (16, 12, 5), (37, 3, 37), (37, 30, 8), (148, 4, 148), (148, 4, 20)])
def test_chain_queue_model_terminates_and_orders_every_slice(G, K, resident):
    for seed in range(6):
        run_chain(G, K, resident, ring=4096, seed=seed)


def test_chain_queue_model_small_ring_is_enough_with_the_slot_handshake():
    # at most G items are pushed and unread at any time: a ring of G + 1 slots never loses one and never blocks
    for seed in range(10):
        run_chain(G=12, K=20, resident=12, ring=13, seed=seed)
        run_chain(G=12, K=20, resident=4, ring=13, seed=seed)
