"""Skeleton topology of the pose-VAE: pooled hierarchies and neighbour masks.

Needed only for the *random-init* path ("random-init weights of the same
architecture", BASELINE.json) -- trained checkpoints carry their masks in the
state dict.  Behavioural reference (studied, not copied):
`python/src/skeleton.py:133-175` (pooling), `:247-271` (which joints collapse),
`:296-362` (ancestor distance, all-pairs distances, neighbour lists) and the
layer plan of `python/src/autoencoder.py:56-110,146-222`.

The one quirk that shapes the masks is kept on purpose: the ancestor-distance
walk gives up as soon as it reaches a child of joint 0, so "j is an ancestor of
i" is only detected when the path does not pass *through* such a child on the
way to j (`skeleton.py:296-304`).  Distances that are not found this way are
recovered by the all-pairs relaxation, so the final matrix is still the tree
metric; the quirk only matters for parents arrays whose root is not 0.
"""
from __future__ import annotations

import numpy as np

INF = float("inf")


def _ancestor_hops(parents, i, j):
    """Hops from i up to ancestor j, or 0 when the walk hits a root child first."""
    hops = 0
    while True:
        if parents[i] == j:
            return hops + 1
        if parents[i] == 0:
            return 0
        i = parents[i]
        hops += 1


def tree_distances(parents):
    n = len(parents)
    d = np.full((n, n), INF)
    np.fill_diagonal(d, 0.0)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            h = _ancestor_hops(parents, i, j)
            if h:
                d[i, j] = d[j, i] = h
    for k in range(n):  # Floyd-Warshall relaxation
        d = np.minimum(d, d[:, k : k + 1] + d[k : k + 1, :])
    return d


def neighbour_lists(parents, max_dist, with_displacement=True):
    """Joints within `max_dist` edges; the displacement pseudo-joint (index J)
    shares the root's neighbourhood and is visible to every root neighbour."""
    d = tree_distances(parents)
    n = len(parents)
    lists = [[j for j in range(n) if d[i, j] <= max_dist] for i in range(n)]
    if with_displacement:
        root_nb = list(lists[0])
        for i in root_nb:
            lists[i].append(n)
        lists.append(root_nb + [n])
    return lists, d


def _collapsing_joints(parents):
    """DFS from the root with a LIFO stack; an inner joint collapses into its
    neighbours when its DFS predecessor has not collapsed (skeleton.py:247-271)."""
    adj, d = neighbour_lists(parents, 1, with_displacement=True)
    n = len(parents)
    degree = [(d[i] == 1).sum() for i in range(n)]
    collapsed, seen = [], set()
    stack = [(0, -1)]
    while stack:
        cur, prev = stack.pop()
        if cur == n:
            continue
        seen.add(cur)
        if prev != -1 and prev not in collapsed and degree[cur] > 1:
            collapsed.append(cur)
        stack.extend((c, cur) for c in adj[cur] if c != cur and c not in seen)
    return collapsed, adj


def pooling_plan(parents, with_displacement=True):
    """Returns (groups, new_parents): groups[i] lists the old joints merged into
    new joint i; with_displacement appends a group covering every old joint."""
    collapsed, adj = _collapsing_joints(parents)
    n = len(parents)
    groups, old2new, new2old = [], {}, {}
    for j in range(n):
        if j not in collapsed:
            old2new[j] = len(groups)
            new2old[len(groups)] = j
            groups.append([j])
    for j in range(n):
        if j in collapsed:
            for nb in adj[j]:
                if nb != j and nb != n:
                    groups[old2new[nb]].append(j)
    new_parents = []
    for i in range(len(groups)):
        p = parents[new2old[i]]
        while p not in old2new:
            p = parents[p]
        new_parents.append(old2new[p])
    if with_displacement:
        groups.append(list(range(n)))
    return groups, new_parents


def _channel_mask(neigh, ch_in, ch_out):
    n = len(neigh)
    m = np.zeros((n * ch_out, n * ch_in), dtype=np.float32)
    for i, nb in enumerate(neigh):
        for k in nb:
            m[i * ch_out : (i + 1) * ch_out, k * ch_in : (k + 1) * ch_in] = 1.0
    return m


def _unpool_matrix(groups, ch):
    outs = set(j for g in groups for j in g)
    n_out = len(outs) + 1
    u = np.zeros((n_out * ch, len(groups) * ch), dtype=np.float32)
    for i, g in enumerate(groups):
        for j in g:
            for c in range(ch):
                u[j * ch + c, i * ch + c] = 1.0
    return u


def _pool_matrix(groups, n_old, ch):
    p = np.zeros((len(groups) * ch, n_old * ch), dtype=np.float32)
    for i, g in enumerate(groups):
        for j in g:
            for c in range(ch):
                p[i * ch + c, j * ch + c] = 1.0 / len(g)
    return p


def decoder_plan(parents, neighbour_distance=2, ch=4, n_layers=3):
    """[(unpool (out,in), mask (out,out))] from the primal skeleton outwards,
    plus the primal feature count (autoencoder.py:146-222)."""
    hier, groups = [list(parents)], []
    cur = list(parents)
    for l in range(n_layers):
        g, cur = pooling_plan(cur, with_displacement=(l != n_layers - 1))
        groups.append(g)
        hier.append(cur)
    layers = []
    for l in range(n_layers):
        lvl = n_layers - l - 1
        neigh, _ = neighbour_lists(hier[lvl], neighbour_distance, with_displacement=True)
        layers.append((_unpool_matrix(groups[lvl], ch), _channel_mask(neigh, ch, ch)))
    return layers, ch * len(hier[-1])


def encoder_plan(parents, neighbour_distance=2, ch=8, n_layers=3):
    """[(mask (n,n), pool (out,in))] from the full skeleton inwards
    (autoencoder.py:56-134)."""
    hier, groups = [list(parents)], []
    cur = list(parents)
    for _ in range(n_layers):
        g, cur = pooling_plan(cur, with_displacement=False)
        groups.append(g)
        hier.append(cur)
    layers = []
    for l in range(n_layers):
        neigh, _ = neighbour_lists(hier[l], neighbour_distance, with_displacement=False)
        layers.append((_channel_mask(neigh, ch, ch), _pool_matrix(groups[l], len(hier[l]), ch)))
    return layers, ch * len(hier[-1])
