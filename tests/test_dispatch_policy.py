"""Host logic of the cluster / streaming dispatch (graph.stream_fwd_preferred, stream_train_preferred):
the measured shapes of profiles/r01_v11_fwd_dispatch_sweep.jsonl must come out on the faster side."""
import types

import pytest

from g_adaptivity_b200 import graph as G


B200_PERSIST_NODES = 148 * 3 * 256      # what gad_wide_persist_nodes answers on a B200 for a 12 KB window


def _fake(n_mesh_nodes, M, C, persist_nodes=B200_PERSIST_NODES):
    g = types.SimpleNamespace(mesh_sizes=[n_mesh_nodes] * M, clf_C=C, cl_C=C, N=n_mesh_nodes * M)
    g.ensure_wide = lambda ce: True
    g.persist_nodes = lambda ce: persist_nodes
    return g


@pytest.mark.parametrize("nodes,M,C,fevals,want_stream", [
    (200 * 200, 1, 16, 256, True),      # cfg 4: 0.41 ms on the persistent kernel, 0.87 ms on one cluster
    (200 * 200, 2, 16, 64, True),       # 0.142 against 0.223 ms (profiles/r02_v3_fwd_dispatch_sweep.jsonl)
    (200 * 200, 4, 16, 64, False),      # beyond the co-resident threads: chain against cluster is a tie -> one launch
    (100 * 100, 1, 16, 64, True),
    (100 * 100, 2, 16, 64, True),
    (100 * 100, 8, 16, 64, True),       # 0.130 against 0.152 ms since the persistent kernel
    (100 * 100, 16, 16, 64, False),     # 160 000 nodes: chain 0.234 against cluster 0.183 ms
    (100 * 100, 64, 4, 64, False),
    (64 * 64, 1, 16, 64, False),
    (64 * 64, 64, 2, 64, False),
    (100 * 100, 1, 16, 4, False),       # four Euler layers: the extra pack launch costs more than it saves
])
def test_forward_dispatch(nodes, M, C, fevals, want_stream, monkeypatch):
    monkeypatch.delenv("GAD_FWD_POLICY", raising=False)
    monkeypatch.delenv("GAD_WIDE_PERSIST", raising=False)
    assert G._stream_fwd_preferred(_fake(nodes, M, C), 4, fevals) == want_stream


def test_forward_dispatch_without_the_persistent_kernel(monkeypatch):
    """GAD_WIDE_PERSIST=0 (or a device where the cooperative launch does not fit): the chain's cost model decides, as
    in round 1 (profiles/r01_v11_fwd_dispatch_sweep.jsonl)."""
    monkeypatch.delenv("GAD_FWD_POLICY", raising=False)
    monkeypatch.setenv("GAD_WIDE_PERSIST", "0")
    assert G._stream_fwd_preferred(_fake(200 * 200, 1, 16), 4, 256)
    assert not G._stream_fwd_preferred(_fake(100 * 100, 8, 16), 4, 64)
    monkeypatch.delenv("GAD_WIDE_PERSIST")
    assert not G._stream_fwd_preferred(_fake(100 * 100, 8, 16, persist_nodes=0), 4, 64)


@pytest.mark.parametrize("nodes,M,C,want_stream", [
    (200 * 200, 1, 16, True),           # 51 us against 70 us
    (100 * 100, 1, 16, False),          # 45 us against 43 us
    (100 * 100, 4, 16, False),
    (100 * 100, 64, 4, False),
    (64 * 64, 3, 16, False),
])
def test_training_dispatch(nodes, M, C, want_stream, monkeypatch):
    monkeypatch.delenv("GAD_TRAIN_POLICY", raising=False)
    assert G._stream_train_preferred(_fake(nodes, M, C), 4) == want_stream


def test_policy_overrides(monkeypatch):
    g = _fake(64 * 64, 1, 16)
    monkeypatch.setenv("GAD_FWD_POLICY", "stream")
    assert G._stream_fwd_preferred(g, 4, 4)
    monkeypatch.setenv("GAD_TRAIN_POLICY", "stream")
    assert G._stream_train_preferred(g, 4)
    g2 = _fake(200 * 200, 1, 16)
    monkeypatch.setenv("GAD_FWD_POLICY", "cluster")
    monkeypatch.setenv("GAD_TRAIN_POLICY", "cluster")
    assert not G._stream_fwd_preferred(g2, 4, 256) and not G._stream_train_preferred(g2, 4)
    g2.ensure_wide = lambda ce: False       # no wide rows (degree > 7): never the chain
    monkeypatch.delenv("GAD_FWD_POLICY")
    assert not G._stream_fwd_preferred(g2, 4, 256)


def test_shared_topology_key_is_host_only_and_shape_and_sample_sensitive():
    """GraphCache.shared_key_of (opt gad_shared_topology): equal for fresh batches on the same mesh, different
    for another batch size or a changed sampled edge / mask entry; reads host tensors only."""
    from g_adaptivity_b200 import synth
    flags = (True, False, 4, "cuda:0")
    a = synth.make_batch((9, 9), 4, seed=1)
    b = synth.make_batch((9, 9), 4, seed=2)          # other features, same mesh
    ka, kb = G.GraphCache.shared_key_of(a, flags), G.GraphCache.shared_key_of(b, flags)
    assert ka == kb and ka[0] == "shared"
    assert G.GraphCache.shared_key_of(synth.make_batch((9, 9), 5, seed=1), flags) != ka
    assert G.GraphCache.shared_key_of(a, flags[:-1] + ("cuda:1",)) != ka
    c = a.clone()
    c.edge_index = c.edge_index.clone()
    c.edge_index[0, 0] += 1                          # column 0 is always sampled
    assert G.GraphCache.shared_key_of(c, flags) != ka
    d = a.clone()
    d.to_boundary_edge_mask = ~d.to_boundary_edge_mask
    assert G.GraphCache.shared_key_of(d, flags) != ka


def test_module_serves_fresh_batches_from_the_shared_topology_key(monkeypatch):
    """GNN._graph with gad_shared_topology: one build, then fresh batches on the same mesh hit the cache
    without the content-fingerprint pass; another batch size builds again.  (Graph build and fingerprint are
    stubbed: this is the host logic only; the GPU run is tests/test_gpu_parity.py.)"""
    import importlib
    import torch
    from g_adaptivity_b200 import synth
    M = importlib.import_module("g_adaptivity_b200.GNN")
    builds, fingerprints = [], []

    def fake_build(*a, **k):
        builds.append(types.SimpleNamespace())
        return builds[-1]

    def fake_content(data, flags, dev, use_masks=True):
        fingerprints.append(1)
        return ("content", len(fingerprints)), {}

    monkeypatch.setattr(M.MeshGraph, "build", staticmethod(fake_build))
    monkeypatch.setattr(G.GraphCache, "content_key_of", staticmethod(fake_content))
    md = (12, 12)
    model = M.GNN(synth.SyntheticDataset(2, md), synth.default_opt(md, gad_shared_topology=True))
    dev = torch.device("cpu")
    gs = [model._graph(synth.make_batch(md, 6, seed=5 + it), dev) for it in range(3)]
    assert len(builds) == 1 and len(fingerprints) == 1 and gs[0] is gs[1] is gs[2]
    assert model._graphs.misses == 1 and model._graphs.shared_hits == 2
    g7 = model._graph(synth.make_batch(md, 7, seed=9), dev)
    assert g7 is not gs[0] and len(builds) == 2 and model._graphs.shared_hits == 2
    # without the option every fresh batch goes through the fingerprint
    model2 = M.GNN(synth.SyntheticDataset(2, md), synth.default_opt(md))
    for it in range(2):
        model2._graph(synth.make_batch(md, 6, seed=5 + it), dev)
    assert len(fingerprints) == 2 + 2 and getattr(model2._graphs, "shared_hits", 0) == 0
