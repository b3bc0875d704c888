"""Host logic of the elimination order of the dense solve (surikatoko_b200/csrc/solve_order.cu) through srk_ba_debug_build_order, no GPU:
whatever the camera graph, the result must be a permutation into [part 0 | ... | separator] with 64-column aligned parts and NO coupling
between two different parts -- the property the concurrent factorisation of the parts relies on."""
import ctypes as C

import numpy as np
import pytest


def build_order(adj, group_size=None):
    import surikatoko_b200 as sb
    L = sb.load_library()
    L.srk_ba_debug_build_order.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    G = adj.shape[0]
    gs = np.full(G, 10, dtype=np.int32) if group_size is None else np.asarray(group_size, dtype=np.int32)
    if group_size is None:
        gs[0], gs[1] = 4, 9                        # quirk Q13: the gauge removes 6 unknowns of frame 0 and 1 of frame 1
    a = np.ascontiguousarray(adj, dtype=np.uint8)
    pos = np.zeros(int(gs.sum()), dtype=np.int32); k0 = np.zeros(32, dtype=np.int32); k1 = np.zeros(32, dtype=np.int32)
    on = C.c_int64(); ksep = C.c_int32()
    nparts = L.srk_ba_debug_build_order(G, gs.ctypes.data, a.ctypes.data, pos.ctypes.data, C.byref(on), k0.ctypes.data, k1.ctypes.data, C.byref(ksep))
    assert nparts >= 0
    L.srk_ba_debug_order_levels.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    m0 = np.zeros(16, dtype=np.int32); m1 = np.zeros(16, dtype=np.int32); msep = C.c_int32()
    nmids = L.srk_ba_debug_order_levels(G, gs.ctypes.data, a.ctypes.data, m0.ctypes.data, m1.ctypes.data, C.byref(msep))
    assert nmids >= 0
    return dict(nparts=nparts, pos=pos, ordered_n=on.value, k0=k0[:nparts], k1=k1[:nparts], ksep=ksep.value, gs=gs,
                nmids=nmids, m0=m0[:nmids], m1=m1[:nmids], msep=msep.value if nparts > 0 else 0)


def check_valid(adj, o):
    gs = o["gs"]; n = int(gs.sum())
    start = np.concatenate([[0], np.cumsum(gs)])
    pos = o["pos"]
    assert len(np.unique(pos)) == n and pos.min() >= 0 and pos.max() < o["ordered_n"]          # injective into the padded range
    if o["nparts"] == 0:
        assert np.array_equal(pos, np.arange(n)) and o["ordered_n"] == n
        return
    tile = pos // 64
    part_of_tile = np.full((o["ordered_n"] + 63) // 64, -1)
    for p, (a, b) in enumerate(zip(o["k0"], o["k1"])):
        assert a < b <= o["ksep"]
        assert np.all(part_of_tile[a:b] == -1)                                              # parts do not overlap
        part_of_tile[a:b] = p
    assert np.all(part_of_tile[:o["ksep"]] >= 0)                                            # every block column before the separator belongs to a part
    part_of_group = np.array([part_of_tile[tile[start[g]]] for g in range(len(gs))])
    for g in range(len(gs)):                                                                # a group sits in one part (or in the separator), contiguous
        assert np.all(part_of_tile[tile[start[g]:start[g + 1]]] == part_of_group[g])
        assert np.array_equal(pos[start[g]:start[g + 1]], pos[start[g]] + np.arange(gs[g]))
    a_idx, b_idx = np.nonzero(adj)
    pa, pb = part_of_group[a_idx], part_of_group[b_idx]
    assert not np.any((pa >= 0) & (pb >= 0) & (pa != pb)), "two different parts are coupled"
    # second level: the separators lie in [ksep, msep), do not overlap, are not coupled to each other, and each touches at most two leaves
    assert o["ksep"] <= o["msep"] <= (o["ordered_n"] + 63) // 64
    mid_of_tile = np.full_like(part_of_tile, -1)
    for m, (a, b) in enumerate(zip(o["m0"], o["m1"])):
        assert o["ksep"] <= a < b <= o["msep"]
        assert np.all(mid_of_tile[a:b] == -1)
        mid_of_tile[a:b] = m
    assert np.all(mid_of_tile[o["ksep"]:o["msep"]] >= 0)
    if o["nmids"] == 0:
        assert o["msep"] == o["ksep"]
        return
    mid_of_group = np.array([mid_of_tile[tile[start[g]]] for g in range(len(gs))])
    ma, mb = mid_of_group[a_idx], mid_of_group[b_idx]
    assert not np.any((ma >= 0) & (mb >= 0) & (ma != mb)), "two second-level separators are coupled"
    for m in range(o["nmids"]):
        touched = set(pb[(ma == m) & (pb >= 0)].tolist()) | set(pa[(mb == m) & (pa >= 0)].tolist())
        assert len(touched) <= 2, "a second-level separator couples more than two leaves"


def ring_graph(M, w, closed=True):
    adj = np.zeros((M, M), dtype=np.uint8)
    for d in range(1, w + 1):
        i = np.arange(M)
        j = i + d
        if closed:
            adj[i, j % M] = 1; adj[j % M, i] = 1
        else:
            ok = j < M
            adj[i[ok], j[ok]] = 1; adj[j[ok], i[ok]] = 1
    return adj


@pytest.mark.parametrize("M,w,closed", [(1000, 9, True), (1000, 9, False), (300, 5, True), (1600, 20, True), (170, 5, True)])
def test_ring_and_chain_graphs_are_split_into_independent_parts(M, w, closed):
    adj = ring_graph(M, w, closed)
    o = build_order(adj)
    check_valid(adj, o)
    assert o["nparts"] >= 4
    nblk0 = (int(o["gs"].sum()) + 63) // 64
    longest = int(np.max(o["k1"] - o["k0"])); sep = (o["ordered_n"] + 63) // 64 - o["msep"]
    longest_mid = int(np.max(o["m1"] - o["m0"])) if o["nmids"] else 0
    assert longest + longest_mid + sep < 0.5 * nblk0, "the dependent chain must be much shorter than in capture order"
    assert o["ordered_n"] < 1.15 * int(o["gs"].sum())                                        # padding stays small


def test_disconnected_graph_two_rings():
    a = ring_graph(400, 6, True)
    adj = np.zeros((800, 800), dtype=np.uint8); adj[:400, :400] = a; adj[400:, 400:] = a
    o = build_order(adj)
    check_valid(adj, o)
    assert o["nparts"] >= 2


def test_grid_graph_street_scene():
    side = 30                                       # 900 cameras on a street grid, each coupled to its 8 neighbours
    idx = np.arange(side * side).reshape(side, side)
    adj = np.zeros((side * side, side * side), dtype=np.uint8)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            a = idx[max(0, dx):side + min(0, dx), max(0, dy):side + min(0, dy)]
            b = idx[max(0, -dx):side + min(0, -dx), max(0, -dy):side + min(0, -dy)]
            adj[a.ravel(), b.ravel()] = 1
    o = build_order(adj)
    check_valid(adj, o)


@pytest.mark.parametrize("kind", ["dense", "star", "small"])
def test_graphs_without_structure_keep_the_natural_order(kind):
    if kind == "dense":
        adj = np.ones((200, 200), dtype=np.uint8)
    elif kind == "star":
        adj = np.zeros((400, 400), dtype=np.uint8); adj[0, :] = 1; adj[:, 0] = 1
    else:
        adj = ring_graph(40, 3, True)               # 393 unknowns: nothing to gain
    o = build_order(adj)
    check_valid(adj, o)
    assert o["nparts"] == 0


def test_random_sparse_graph_is_still_valid():
    rng = np.random.default_rng(3)
    M = 500
    adj = ring_graph(M, 4, False)
    extra = rng.integers(0, M, size=(60, 2))        # a few long-range co-visibilities (loop closures)
    adj[extra[:, 0], extra[:, 1]] = 1; adj[extra[:, 1], extra[:, 0]] = 1
    o = build_order(adj)
    check_valid(adj, o)
