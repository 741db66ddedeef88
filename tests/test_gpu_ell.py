"""GPU parity of the mesh-resident ELL kernels (csrc/ell_kernels.cuh) and of the one-launch training
pass `gad_deform_train_ell`, against the CPU oracle and against the CSR mesh-resident kernels.

Bars as in test_gpu_parity.py: coordinates 1e-5 relative, parameter gradients 1e-4 relative.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import gad_testutil as util
from g_adaptivity_b200 import GNN, synth
from g_adaptivity_b200.trainer import DeformerTrainer
from oracle import gnn_oracle
from test_gpu_parity import COORD_TOL, GRAD_TOL, check_grads, cuda_model, grads_of, oracle_model

pytestmark = pytest.mark.gpu


def _case(mesh_dims, B, seed=0, burgers=False, **over):
    opt = synth.burgers_opt(mesh_dims, **over) if burgers else synth.default_opt(mesh_dims, **over)
    ds = synth.SyntheticDataset(len(mesh_dims), mesh_dims)
    data = synth.make_batch(mesh_dims, B, seed=seed, burgers=burgers)
    torch.manual_seed(42)
    ref = oracle_model(ds, opt)
    return opt, ds, data, ref


@pytest.mark.parametrize("mesh_dims,B,burgers,over,slots", [
    ((30, 30), 16, False, {}, 6),                       # cfg2 shape: CE=4, 6 slots
    ((12, 12), 5, False, {"self_loops": True}, 7),      # +self-loops: 7 slots
    ((200,), 64, True, {}, 2),                          # 1-D Burgers: CE=2, 2 slots
    ((40,), 9, True, {"self_loops": True}, 3),          # 1-D + self-loops: 3 slots
    ((10, 10), 4, False, {"gnn_inc_feat_f": False, "gnn_inc_feat_uu": False}, 6),   # 2-D, CE=2
    ((50, 50), 4, False, {}, 6),                        # cfg5 shape: ELL rows streamed from L2
    ((9, 9), 3, False, {"fix_boundary": False}, 6),
])
def test_ell_path_matches_oracle_and_csr_path(mesh_dims, B, burgers, over, slots):
    opt, ds, data, ref = _case(mesh_dims, B, burgers=burgers, **over)
    ref_out = ref(data)
    gnn_oracle.mesh_loss(ref_out, data.x_phys).backward()
    ref_grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    g64, floor, scale = util.fp64_grads_and_noise_floor(ds, opt, data, ref, ref_grads)
    outs = {}
    for no_ell in (False, True):
        model = cuda_model(ds, opt, ref.state_dict(), gad_no_ell=no_ell)
        model.train()
        out = model(data)
        g = model.last_graph
        assert g.tile_ptr is not None
        assert (g.ell_in is None) == no_ell
        if not no_ell:
            assert max(g.max_in_deg, g.max_out_deg) <= slots
        assert util.rel_err(out, ref_out) <= COORD_TOL
        tgt = data.x_phys.cuda()
        F.l1_loss(out, tgt if tgt.dim() == 2 else tgt.unsqueeze(-1)).backward()
        util.check_grads_conditioned(grads_of(model), g64, floor, scale)
        outs[no_ell] = out.detach().cpu()
    assert util.rel_err(outs[False], outs[True]) <= 2e-6


def test_ell_rows_encode_the_csr():
    """ELL rows are a re-encoding of the CSR / CSC arrays, integer-exact: the valid slots of row i
    hold exactly the neighbours of i (as a multiset), the other slots point at rows inside the tile;
    on a structured mesh every slot means one neighbour offset (conflict-free gathers)."""
    opt, ds, data, ref = _case((13, 13), 6)
    model = cuda_model(ds, opt, ref.state_dict())
    model(data)
    g = model.last_graph
    tp = g.tile_ptr.cpu().numpy()
    tile_of = np.searchsorted(tp, np.arange(g.N), side="right") - 1
    rb = g.ell_ce * 4
    for ell, ptr, idx in ((g.ell_in, g.rowptr, g.col_walk), (g.ell_out, g.t_rowptr, g.t_dst_walk)):
        rows = ell.cpu().numpy().view(np.uint16).astype(np.int64)
        ptr, idx = ptr.cpu().numpy(), idx.cpu().numpy()
        offsets_of_slot = [set() for _ in range(7)]
        for i in range(g.N):
            n0, n1 = tp[tile_of[i]], tp[tile_of[i] + 1]
            valid = [q for q in range(7) if (rows[i, 7] >> q) & 1]
            assert all(q < 6 for q in valid)                              # max degree 6 -> 6 slots in use
            got = sorted(rows[i, q] // rb + n0 for q in valid)
            assert got == sorted(idx[ptr[i]:ptr[i + 1]].tolist())
            assert all(rows[i, q] % rb == 0 and n0 <= rows[i, q] // rb + n0 < n1 for q in range(7))
            if len(valid) == 6:
                for q in valid:
                    offsets_of_slot[q].add(int(rows[i, q] // rb + n0 - i))
        assert all(len(sset) <= 1 for sset in offsets_of_slot)            # one offset per slot (interior nodes)


@pytest.mark.parametrize("loss", ["l1", "mse"])
@pytest.mark.parametrize("over", [
    {},
    {"share_conv": False},
    {"learn_step": True},
    {"share_conv": False, "learn_step": True, "num_layers": 3},
    {"num_layers": 1},
    {"softmax_temp_type": "fixed", "softmax_temp": 2.0},
])
def test_fused_train_step_matches_oracle(loss, over):
    """One launch per tile for pack + forward + loss + backward: loss value and every parameter
    gradient against autograd on the oracle."""
    opt, ds, data, ref = _case((14, 14), 11, seed=4, loss_fn=loss, **over)
    ref_out = ref(data)
    ref_loss = gnn_oracle.mesh_loss(ref_out, data.x_phys, loss_fn=loss)
    ref_loss.backward()
    ref_grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    model = cuda_model(ds, opt, ref.state_dict())
    tr = DeformerTrainer(model, use_cuda_graph=False, loss_fn=loss)
    sid = tr.add_batch(data)
    assert tr.slots[sid].graph.ell_in is not None
    with torch.cuda.stream(tr.stream):
        tr._issue(tr.slots[sid], tr.stream.cuda_stream, with_optimizer=False)
    tr.synchronize()
    assert abs(tr.slots[sid].loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert util.rel_err(tr.slots[sid].x_phys, ref_out) <= COORD_TOL
    check_grads(grads_of(model), ref_grads)


def test_fused_train_equals_unfused_kernels():
    opt, ds, data, ref = _case((30, 30), 24, seed=6)
    res = []
    for no_fused in (False, True):
        model = cuda_model(ds, opt, ref.state_dict(), gad_no_fused_train=no_fused)
        tr = DeformerTrainer(model, use_cuda_graph=False)
        sid = tr.add_batch(data)
        with torch.cuda.stream(tr.stream):
            tr._issue(tr.slots[sid], tr.stream.cuda_stream, with_optimizer=False)
        tr.synchronize()
        res.append((tr.slots[sid].loss.item(), tr.gflat.clone().cpu()))
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[1][0])
    scale = res[1][1].abs().max().item()
    assert (res[0][1] - res[1][1]).abs().max().item() <= 1e-5 * scale


def test_fused_train_graph_replay_is_deterministic_and_trains():
    opt, ds, data, ref = _case((30, 30), 32, seed=8)
    finals = []
    for _ in range(2):
        model = cuda_model(ds, opt, ref.state_dict())
        tr = DeformerTrainer(model, lr=1e-2)
        sid = tr.add_batch(data)
        losses = []
        for _ in range(12):
            loss = tr.step(sid)
            with torch.cuda.stream(tr.stream):      # the step runs on the trainer's stream
                losses.append(loss.clone())
        tr.synchronize()
        losses = [float(x.item()) for x in losses]
        assert losses[-1] < losses[0]
        finals.append((losses, tr.flat.clone().cpu()))
    assert finals[0][0] == finals[1][0]
    assert torch.equal(finals[0][1], finals[1][1])


@pytest.mark.parametrize("use_graph", [False, True])
def test_one_launch_step_equals_multi_kernel_step(use_graph):
    """Adam + refold in the tail of the train kernel == gad_weight_grads / gad_adam_step /
    gad_prepare_weights as separate launches: same parameters after several steps."""
    opt, ds, data, ref = _case((20, 20), 12, seed=5, learn_step=True)
    res = []
    for no_fused in (False, True):
        model = cuda_model(ds, opt, ref.state_dict(), gad_no_fused_train=no_fused)
        tr = DeformerTrainer(model, lr=3e-3, use_cuda_graph=use_graph)
        sid = tr.add_batch(data)
        for _ in range(6):
            tr.step(sid)
        tr.synchronize()
        res.append((tr.flat.clone().cpu(), tr.slots[sid].loss.item(), int(tr.step_count.item())))
    assert res[0][2] == res[1][2] == 6
    assert abs(res[0][1] - res[1][1]) <= 2e-5 * abs(res[1][1])
    assert (res[0][0] - res[1][0]).abs().max().item() <= 2e-5 * res[1][0].abs().max().item()
    # the model's parameters are views of the trainer's flat vector: state_dict sees the update
    sd = model.state_dict()
    assert not torch.equal(sd["conv_layers.0.lin_query.weight"].cpu(), ref.state_dict()["conv_layers.0.lin_query.weight"])


@pytest.mark.parametrize("pdl", [True, False])
def test_epoch_graph_equals_step_by_step(pdl):
    """One CUDA graph for a pass over the ring (programmatic dependent launch between the steps)
    trains exactly like one graph replay per step."""
    opt, ds, _, ref = _case((30, 30), 8, seed=1)
    batches = [synth.make_batch((30, 30), 40, seed=20 + r) for r in range(3)]
    res = []
    for epoch_mode in (True, False):
        model = cuda_model(ds, opt, ref.state_dict(), gad_pdl=pdl)
        tr = DeformerTrainer(model, lr=5e-3)
        sids = [tr.add_batch(b) for b in batches]
        for _ in range(4):
            if epoch_mode:
                tr.run_epoch(sids)
            else:
                for sid in sids:
                    tr.step(sid)
        tr.synchronize()
        res.append((tr.flat.clone().cpu(), [tr.slots[s].loss.item() for s in sids], int(tr.step_count.item())))
    assert res[0][2] == res[1][2] == 12
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1]


# ------------------------------------------------------------------------------------------
# cluster-resident kernel (csrc/cl_kernels.cu): meshes beyond one CTA, one thread-block cluster each
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mesh_dims,B,over", [
    ((64, 64), 3, {}),                                        # 4096 nodes: cluster of 2
    ((81, 81), 2, {"loss_fn": "mse"}),                         # 6561 nodes: cluster of 4, ragged last slab
    ((64, 64), 2, {"share_conv": False, "learn_step": True, "num_layers": 3}),
    ((64, 64), 2, {"learn_step": True}),                       # shared weights + step-size partials
])
def test_cluster_train_step_matches_oracle_and_streaming_path(mesh_dims, B, over):
    loss = over.get("loss_fn", "l1")
    opt, ds, data, ref = _case(mesh_dims, B, seed=9, **over)
    ref_out = ref(data)
    ref_loss = gnn_oracle.mesh_loss(ref_out, data.x_phys, loss_fn=loss)
    ref_loss.backward()
    ref_grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    res = []
    for no_cluster in (False, True):
        model = cuda_model(ds, opt, ref.state_dict(), gad_no_cluster=no_cluster)
        tr = DeformerTrainer(model, use_cuda_graph=False, loss_fn=loss)
        sid = tr.add_batch(data)
        g = tr.slots[sid].graph
        assert g.tile_ptr is None and (g.cl_in is not None) == (not no_cluster)
        with torch.cuda.stream(tr.stream):
            tr._issue(tr.slots[sid], tr.stream.cuda_stream, with_optimizer=False)
        tr.synchronize()
        assert abs(tr.slots[sid].loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
        assert util.rel_err(tr.slots[sid].x_phys, ref_out) <= COORD_TOL
        res.append(tr.gflat.clone().cpu())
        if not no_cluster:
            assert g.cl_C >= 2
            g64, floor, scale = util.fp64_grads_and_noise_floor(ds, opt, data, ref, ref_grads)
            util.check_grads_conditioned(grads_of(model), g64, floor, scale)
    scale = res[1].abs().max().item()
    assert (res[0] - res[1]).abs().max().item() <= 2e-5 * scale      # cluster kernel vs streaming ELL kernels


def test_cluster_training_is_deterministic_and_matches_streaming_steps():
    """Several Adam steps (tail of the cluster kernel, PDL-chained graph replay) == the same steps on
    the streaming kernels with stand-alone weight-gradient / Adam / refold launches."""
    opt, ds, data, ref = _case((64, 64), 3, seed=10)
    finals = []
    for mode in ("cluster", "cluster", "stream"):
        model = cuda_model(ds, opt, ref.state_dict(), gad_no_cluster=(mode == "stream"))
        tr = DeformerTrainer(model, lr=1e-2)
        sids = [tr.add_batch(data), tr.add_batch(data)]
        losses = []
        for _ in range(4):
            for l in tr.run_epoch(sids):
                with torch.cuda.stream(tr.stream):
                    losses.append(l.clone())
        tr.synchronize()
        finals.append(([float(x.item()) for x in losses], tr.flat.clone().cpu()))
    assert finals[0][0] == finals[1][0] and torch.equal(finals[0][1], finals[1][1])     # bit-reproducible
    assert finals[0][0][-1] < finals[0][0][0]
    assert (finals[0][1] - finals[2][1]).abs().max().item() <= 2e-4 * finals[2][1].abs().max().item()


def test_single_very_large_mesh_trains_on_the_streaming_chain(monkeypatch):
    """One 128x128 mesh would keep a 16-CTA cluster on 16 SMs: the trainer picks the chain of dependent
    streaming launches (graph.stream_train_preferred); forced onto the cluster kernel it takes the same steps."""
    opt, ds, data, ref = _case((128, 128), 1, seed=13)
    finals = []
    for policy in (None, "cluster"):
        if policy:
            monkeypatch.setenv("GAD_TRAIN_POLICY", policy)
        model = cuda_model(ds, opt, ref.state_dict())
        tr = DeformerTrainer(model, lr=1e-2)
        sid = tr.add_batch(data)
        s = tr.slots[sid]
        assert s.graph.cl_in is not None and s.graph.cl_C == 16
        assert tr._cluster(s) == (policy == "cluster")
        losses = []
        for _ in range(3):
            l = tr.step(sid)
            with torch.cuda.stream(tr.stream):
                losses.append(l.clone())
        tr.synchronize()
        finals.append(([float(x.item()) for x in losses], tr.flat.clone().cpu()))
    assert finals[0][0][-1] < finals[0][0][0]
    assert abs(finals[0][0][0] - finals[1][0][0]) <= 1e-5 * abs(finals[1][0][0])
    assert (finals[0][1] - finals[1][1]).abs().max().item() <= 2e-4 * finals[1][1].abs().max().item()


def test_host_fed_loop_packed_buffers_equal_per_tensor_copies():
    """run_from_host over pack_host buffers (one H2D copy per step; loop in the C library, gad_pipeline_run, or
    the same schedule from Python; whole buffer or the per-sample prefix without the shared x_comp) == over
    Batch objects (one copy per tensor) == resident-batch steps: same losses, same parameters."""
    opt, ds, _, ref = _case((20, 20), 16, seed=11)
    batches = [synth.make_batch((20, 20), 16, seed=20 + r) for r in range(3)]
    for b in batches:
        b.pin_memory()
    outs = []
    for mode in ("resident", "batches", "packed_native", "packed_python", "prefix_native"):
        model = cuda_model(ds, opt, ref.state_dict())
        tr = DeformerTrainer(model, lr=1e-2)
        sids = [tr.add_batch(b) for b in batches]
        if mode == "resident":
            losses = []
            for k in range(9):
                l = tr.step(sids[k % 3])
                with torch.cuda.stream(tr.stream):
                    losses.append(l.clone())
            tr.synchronize()
            losses = torch.stack([x.cpu() for x in losses]).reshape(-1)
        elif mode == "batches":
            losses = tr.run_from_host(batches, 9).clone()
        else:
            src = [tr.pack_host(sid, b, with_x_comp=not mode.startswith("prefix")) for sid, b in zip(sids, batches)]
            if mode.startswith("prefix"):
                assert src[0].numel() * 4 == 16 * 400 * 16          # target (8) + f (4) + uu (4) bytes per node
            losses = tr.run_from_host(src, 9, native=mode.endswith("native")).clone()
            assert tr.slots[0].h2d_bytes == src[0].numel() * 4
        outs.append((losses, tr.flat.clone().cpu()))
    for losses, flat in outs[1:]:
        assert torch.equal(losses, outs[0][0])
        assert torch.equal(flat, outs[0][1])
    # a batch on another mesh cannot be sent without its x_comp
    other = synth.make_batch((20, 20), 16, seed=99)
    other.x_comp = other.x_comp + 0.5
    with pytest.raises(ValueError):
        tr.pack_host(0, other, with_x_comp=False)


@pytest.mark.parametrize("fraction", [0.3, 0.77])
def test_host_fed_loop_with_a_relay_through_another_gpu(fraction):
    """run_from_host(relay=(device, fraction)): part of every batch reaches the GPU through ANOTHER GPU's path to host
    memory (host -> staging there -> NVLink, gad_pipeline_run_relay).  Same losses, same parameters as the direct
    loop, bit for bit.  Needs two visible GPUs (the driver's single-GPU lease skips it)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    opt, ds, _, ref = _case((20, 20), 16, seed=11)
    batches = [synth.make_batch((20, 20), 16, seed=20 + r) for r in range(3)]
    outs = []
    for relay in (None, (1, fraction)):
        model = cuda_model(ds, opt, ref.state_dict())
        tr = DeformerTrainer(model, lr=1e-2)
        sids = [tr.add_batch(b) for b in batches]
        src = [tr.pack_host(sid, b, with_x_comp=False) for sid, b in zip(sids, batches)]
        losses = tr.run_from_host(src, 11, relay=relay).clone()
        again = tr.run_from_host(src, 5, relay=relay).clone()        # staging and events are reused
        outs.append((losses, again, tr.flat.clone().cpu()))
        tr.close()
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        tr2 = DeformerTrainer(cuda_model(ds, opt, ref.state_dict()), lr=1e-2)
        s2 = [tr2.add_batch(b) for b in batches]
        tr2.run_from_host([tr2.pack_host(i, b) for i, b in zip(s2, batches)], 3, relay=(0, 0.5))   # itself


# ------------------------------------------------------------------------------------------
# shared-topology batches (row f2): one ELL table per mesh shape, graph build O(mesh)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mesh_dims,B,burgers", [((30, 30), 9, False), ((12, 12), 37, False), ((50, 50), 3, False),
                                                 ((40,), 70, True)])
def test_shared_topology_graph_equals_general_graph(mesh_dims, B, burgers):
    """`opt['gad_shared_topology']` (dataset on one mesh, src/data.py:143): the graph prologue runs for the meshes
    of ONE tile and every tile uses that tile's ELL table (tile_ptr = NULL in the C ABI).  Module forward +
    backward and K training steps must equal the general graph's bit for bit -- same kernels, same rows -- with
    topology tensors the size of one tile."""
    opt, ds, data, ref = _case(mesh_dims, B, seed=13, burgers=burgers)
    N1 = int(np.prod(mesh_dims))
    outs = []
    for shared in (False, True):
        model = cuda_model(ds, opt, ref.state_dict(), gad_shared_topology=shared, gad_store_alpha=False)
        model.train()
        out = model(data)
        g = model.last_graph
        assert g.uniform == shared
        tgt = data.x_phys.cuda()
        F.l1_loss(out, tgt if tgt.dim() == 2 else tgt.unsqueeze(-1)).backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        with torch.no_grad():
            out_inf = model(data)
        tr = DeformerTrainer(model, lr=1e-2, loss_fn="l1")
        sid = tr.add_batch(data)
        assert tr.slots[sid].graph.uniform == shared
        for _ in range(5):
            tr.step(sid)
        tr.synchronize()
        outs.append((out.detach().clone(), grads, out_inf.clone(), tr.flat.clone(), tr.slots[sid].loss.item(), g))
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])
    for n in a[1]:
        assert torch.equal(a[1][n], b[1][n]), n
    assert torch.equal(a[3], b[3]) and a[4] == b[4]
    g_gen, g_sh = a[5], b[5]
    assert g_sh.E == g_gen.E and g_sh.N == g_gen.N and g_sh.T == g_gen.T
    assert g_sh.max_tile_nodes == g_gen.max_tile_nodes
    assert g_sh.ell_in.shape[0] == g_sh.max_tile_nodes < g_gen.ell_in.shape[0] or B * N1 == g_sh.max_tile_nodes
    assert torch.equal(g_sh.ell_in, g_gen.ell_in[:g_sh.max_tile_nodes])          # tile-local rows: tile 0's table
    assert torch.equal(g_sh.tile_ptr, g_gen.tile_ptr)


def test_shared_topology_falls_back_when_the_batch_breaks_the_promise():
    """Mixed mesh sizes, or a batch whose last mesh carries a different edge block, get the general graph."""
    opt, ds, data, ref = _case((9, 9), 6, seed=2)
    model = cuda_model(ds, opt, ref.state_dict(), gad_shared_topology=True, gad_store_alpha=False)
    bad = data.clone()
    E1 = bad.edge_index.shape[1] // 6
    perm = torch.randperm(E1, generator=torch.Generator().manual_seed(0))
    for name in ("to_boundary_edge_mask", "to_corner_nodes_mask", "diff_boundary_edges_mask"):
        t = getattr(bad, name).clone()
        t[5 * E1:] = t[5 * E1:][perm]
        setattr(bad, name, t)
    ei = bad.edge_index.clone()
    ei[:, 5 * E1:] = ei[:, 5 * E1:][:, perm]
    bad.edge_index = ei
    with torch.no_grad():
        out_bad = model(bad)
        assert not model.last_graph.uniform
        out = model(data)
        assert model.last_graph.uniform
    assert util.rel_err(out_bad, out) <= 2e-6          # same graph, edges listed in another order
    mixed = synth.make_mixed_batch([(5, 5), (7, 7), (6, 6)], seed=1)
    m2 = cuda_model(synth.SyntheticDataset(2, (5, 5)), synth.default_opt((5, 5)), gad_shared_topology=True,
                    gad_store_alpha=False)
    with torch.no_grad():
        m2(mixed)
    assert not m2.last_graph.uniform


@pytest.mark.parametrize("shared", [True, False])
def test_train_batch_keeps_the_reference_loop_shape(shared):
    """`for data in loader: loss = trainer.train_batch(data)` -- a fresh host Batch per iteration, slots found by
    topology -- trains exactly like stepping over resident batches: same losses, same parameters."""
    opt, ds, _, ref = _case((20, 20), 12, seed=3)
    fresh = [synth.make_batch((20, 20), 12, seed=50 + k) for k in range(7)]
    for b in fresh:
        b.pin_memory()
    model = cuda_model(ds, opt, ref.state_dict(), gad_shared_topology=shared, gad_store_alpha=False)
    tr = DeformerTrainer(model, lr=1e-2)
    got = []
    for b in fresh:
        loss = tr.train_batch(b if shared else b)
        tr.synchronize()
        got.append(float(loss.item()))
    n_slots = len(tr.slots)
    assert n_slots == (3 if shared else 7)                 # a shared mesh: three slots serve every batch
    model2 = cuda_model(ds, opt, ref.state_dict(), gad_store_alpha=False)
    tr2 = DeformerTrainer(model2, lr=1e-2)
    want = []
    for b in fresh:
        sid = tr2.add_batch(b)
        l = tr2.step(sid)
        tr2.synchronize()
        want.append(float(l.item()))
    assert got == want
    assert torch.equal(tr.flat, tr2.flat)
