"""Exhaustive interleaving check of the peer-memory exchange protocol (3dssd_b200/csrc/peer_gather.cu).

The kernel's safety rests on an argument in its header: two parities of receive area + flags make re-use safe without
acknowledgements.  This test does not trust the argument: it explores EVERY interleaving of the protocol's steps for small
worlds and a few replays and checks that each copy-out reads exactly the slice its peer sent in that replay.  The same
explorer finds the violation when the protocol is weakened (one parity; flag published before the data), so a green
result is not vacuous.  A rank's replays are sequential (same stream: replay s+1 starts when every CTA of replay s has
finished), CTA p of a replay runs   W: slice -> peer p's recv[par][me]   F: peer p's flag[par][me] = s
wait: my flag[par][p] >= s   R: copy my recv[par][p] out -- the four program points of the kernel.  There is no
multi-GPU sharding in the reference to compare with (lib/core/evaluator.py:145-147 is single-GPU); the GPU test
test_peer_allgather_protocol_three_simulated_ranks runs the real kernel."""

import pytest

W_, F_, WAIT_, R_ = range(4)


def explore(world, replays, parities=2, flag_first=False, self_cta=True):
    """Depth-first search over all interleavings.  Returns None if every reachable copy-out is correct and every run
    terminates, else a description of the first violation.  self_cta=False leaves out CTA `rank` of every rank: it touches
    only slot [rank][parity][rank] of its own memory, which no other CTA reads or writes, so it cannot take part in a
    violation -- dropping it keeps the 3-rank search small."""
    order = (F_, W_, WAIT_, R_) if flag_first else (W_, F_, WAIT_, R_)

    # state: (seq per rank, pc index per (rank, cta), recv tags per (rank, parity, src), flags per (rank, parity, src))
    def key(seq, pc, recv, flag):
        return (tuple(seq), tuple(pc), tuple(recv), tuple(flag))

    def ix(r, par, src):
        return (r * parities + par) * world + src

    start = ([1] * world, [0 if (self_cta or q // world != q % world) else 4 for q in range(world * world)], [0] * (world * parities * world), [0] * (world * parities * world))
    seen = set()
    stack = [start]
    while stack:
        seq, pc, recv, flag = stack.pop()
        k = key(seq, pc, recv, flag)
        if k in seen:
            continue
        seen.add(k)
        moved = False
        finished = all(s > replays for s in seq)
        for r in range(world):
            s = seq[r]
            if s > replays:
                continue
            par = s % parities
            for p in range(world):
                i = pc[r * world + p]
                if i == 4:
                    continue
                op = order[i]
                nseq, npc, nrecv, nflag = seq, list(pc), recv, flag
                if op == W_:
                    nrecv = list(recv); nrecv[ix(p, par, r)] = s            # tag = replay number of the sender's slice
                elif op == F_:
                    nflag = list(flag); nflag[ix(p, par, r)] = s
                elif op == WAIT_:
                    if flag[ix(r, par, p)] < s:
                        continue                                             # still spinning
                else:
                    if recv[ix(r, par, p)] != s:
                        return "rank %d replay %d read peer %d's slice of replay %d" % (r, s, p, recv[ix(r, par, p)])
                npc[r * world + p] = i + 1
                if all(npc[r * world + q] == 4 for q in range(world)):       # kernel of replay s complete: next replay may start
                    nseq = list(seq); nseq[r] = s + 1
                    for q in range(world):
                        npc[r * world + q] = 0 if (self_cta or q != r) else 4
                stack.append((nseq, npc, nrecv, nflag))
                moved = True
        if not moved and not finished:
            return "deadlock at replays %s" % (seq,)
    return None


@pytest.mark.parametrize("world,replays,self_cta", [(2, 6, True), (3, 5, False)])
def test_two_parity_protocol_is_safe_under_every_interleaving(world, replays, self_cta):
    assert explore(world, replays, self_cta=self_cta) is None


def test_single_parity_protocol_is_caught():
    """With one receive area a fast peer overwrites a slice that has not been copied out yet: the explorer must find it."""
    assert explore(2, 3, parities=1) is not None


def test_flag_before_data_is_caught():
    """Publishing the flag before the slice is written lets the peer copy stale data: the explorer must find it."""
    assert explore(2, 2, flag_first=True) is not None
