"""GPU parity tests, kernel by kernel, through the C-ABI (ops.py -> libignnition_b200.so) against the
CPU oracle on the same seeded inputs.  Integer outputs are bit-exact; float outputs are compared
with the north-star tolerance (1e-5 relative, fp32)."""

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ignnition_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5          # north-star tolerance (BASELINE.json): 1e-5 relative, fp32


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def rel_err(got, want):
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def random_edges(rng, n_dst, n_src, max_len, shuffle=True, empty_frac=0.2):
    lens = rng.randint(0, max_len + 1, n_dst)
    lens[rng.rand(n_dst) < empty_frac] = 0
    dst = np.repeat(np.arange(n_dst), lens)
    seq = np.concatenate([np.arange(l) for l in lens]) if lens.sum() else np.zeros(0, np.int64)
    src = rng.randint(0, n_src, len(dst))
    if shuffle:      # groups in arbitrary order, seq ascending inside each destination (reference :145-153)
        order = np.argsort(rng.permutation(n_dst)[dst], kind="stable")
        dst, seq, src = dst[order], seq[order], src[order]
    return src, dst, seq


# ------------------------------------------------------------------ CSR builder (bit-exact)
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n_dst,n_src,max_len", [(1, 1, 1), (7, 5, 3), (1000, 300, 9), (70000, 5000, 40), (5, 3, 0)])
def test_csr_build_bit_exact(mode, n_dst, n_src, max_len):
    from ignnition_b200 import ops
    rng = np.random.RandomState(n_dst + max_len)
    src, dst, seq = random_edges(rng, n_dst, n_src, max_len)
    rowptr, col, perm, status = ops.csr_build(dev(dst, torch.int32), dev(src, torch.int32), dev(seq, torch.int32),
                                              n_dst, mode, want_perm=True, want_status=True)
    r, c, p = orc.csr_from_edges(src, dst, seq, n_dst)
    assert np.array_equal(rowptr.cpu().numpy(), r)
    assert np.array_equal(col.cpu().numpy(), c)
    assert np.array_equal(perm.cpu().numpy(), p)
    st = status.cpu().numpy()
    assert st[0] == 0 and st[1] == (np.diff(r).max() if n_dst else 0)


def test_csr_build_sort_without_seq_and_large():
    from ignnition_b200 import ops
    rng = np.random.RandomState(3)
    n_dst, E = 3_000_000, 5_000_000          # 22-bit keys: 3 radix passes, multi-level scan
    dst = rng.randint(0, n_dst, E)
    src = rng.randint(0, 1 << 20, E)
    rowptr, col, perm, _ = ops.csr_build(dev(dst, torch.int32), dev(src, torch.int32), None, n_dst, 0, want_perm=True)
    r, c, p = orc.stable_sort_csr(src, dst, n_dst)
    assert np.array_equal(rowptr.cpu().numpy(), r)
    assert np.array_equal(perm.cpu().numpy(), p)
    assert np.array_equal(col.cpu().numpy(), c)


def test_csr_build_golden_and_status_flags_bad_seq(golden):
    from ignnition_b200 import ops
    g = golden("routenet_geant2")
    ref = g["reference_tensors"][0]
    for adj, s_e, d_e, _ in g["reference_meta"]["adjacency_info"]:
        src, dst, seq = (np.array(ref["src_" + adj]), np.array(ref["dst_" + adj]),
                         np.array(ref["seq_%s_%s" % (s_e, d_e)]))
        n = ref["num_" + d_e]
        for mode in (0, 1):
            rowptr, col, perm, st = ops.csr_build(dev(dst, torch.int32), dev(src, torch.int32),
                                                  dev(seq, torch.int32), n, mode, True, True)
            r, c, p = orc.csr_from_edges(src, dst, seq, n)
            assert np.array_equal(rowptr.cpu().numpy(), r) and np.array_equal(col.cpu().numpy(), c)
            assert np.array_equal(perm.cpu().numpy(), p) and st.cpu().numpy()[0] == 0
        bad = seq.copy()
        bad[0], bad[1] = bad[1], bad[0]
        if bad[0] != seq[0]:
            _, _, _, st = ops.csr_build(dev(dst, torch.int32), dev(src, torch.int32), dev(bad, torch.int32), n, 0,
                                        True, True)
            assert st.cpu().numpy()[0] > 0


def test_length_order_is_stable_descending():
    from ignnition_b200 import ops
    rng = np.random.RandomState(5)
    lens = rng.randint(0, 9, 100_000)
    rowptr = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=rowptr[1:])
    order = ops.length_order(dev(rowptr, torch.int32)).cpu().numpy()
    want = np.argsort(-lens, kind="stable")
    assert np.array_equal(order, want)


# ------------------------------------------------------------------ gather + segmented aggregation
@pytest.mark.parametrize("F", [4, 8, 32, 64, 100, 256])
@pytest.mark.parametrize("op", [0, 1, 2])
def test_segment_reduce(F, op):
    from ignnition_b200 import ops
    rng = np.random.RandomState(F + op)
    n_dst, n_src = 3001, 777
    src, dst, seq = random_edges(rng, n_dst, n_src, 37)
    states = rng.randn(n_src, F).astype(np.float32)
    r, c, _ = orc.csr_from_edges(src, dst, seq, n_dst)
    got = ops.segment_reduce(op, dev(r, torch.int32), dev(c, torch.int32), dev(states)).cpu().numpy()
    want = np.zeros((n_dst, F), np.float64)
    for d in range(n_dst):
        rows = states[c[r[d]:r[d + 1]]].astype(np.float64)
        if len(rows):
            want[d] = rows.sum(0) if op == 0 else rows.mean(0) if op == 1 else rows.max(0)
    if op == 2:
        assert np.array_equal(got, want.astype(np.float32))          # max is exact
    else:
        assert rel_err(got, want) < RTOL
    if op == 0:   # slot-order single-accumulator sum == sequential fp32 sum, bit for bit
        seqsum = np.zeros((n_dst, F), np.float32)
        for d in range(n_dst):
            acc = np.zeros(F, np.float32)
            for e in range(r[d], r[d + 1]):
                acc = acc + states[c[e]]
            seqsum[d] = acc
        assert np.array_equal(got, seqsum)


def test_segment_reduce_add_accumulates():
    """IGN_OP_SUM_ADD: out += segment sum (partial sums over edge buckets); two halves of an edge list == the whole"""
    from ignnition_b200 import ops
    rng = np.random.RandomState(11)
    n_dst, n_src, F = 3000, 800, 64
    src, dst, seq = random_edges(rng, n_dst, n_src, 9)
    states = rng.randn(n_src, F).astype(np.float32)
    half = rng.rand(len(dst)) < 0.5
    outs = []
    for sel in (half, ~half):
        r, c, _ = orc.stable_sort_csr(src[sel], dst[sel], n_dst)
        outs.append((dev(r, torch.int32), dev(c, torch.int32)))
    out = ops.segment_reduce(ops.OP_SUM, outs[0][0], outs[0][1], dev(states))
    ops.segment_reduce(ops.OP_SUM_ADD, outs[1][0], outs[1][1], dev(states), out=out)
    want = np.zeros((n_dst, F), np.float64)
    np.add.at(want, dst, states[src].astype(np.float64))
    assert rel_err(out.cpu().numpy(), want) < RTOL


def test_segment_reduce_identity_col_and_errors():
    from ignnition_b200 import ops
    rng = np.random.RandomState(9)
    msgs = rng.randn(50, 32).astype(np.float32)
    rowptr = np.array([0, 10, 10, 50], np.int32)
    got = ops.segment_reduce(0, dev(rowptr), None, dev(msgs)).cpu().numpy()
    assert rel_err(got[0], msgs[:10].sum(0)) < RTOL and np.all(got[1] == 0)
    with pytest.raises(RuntimeError, match="IGNNITION"):
        ops.segment_reduce(0, dev(rowptr), None, dev(rng.randn(50, 30).astype(np.float32)))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.segment_reduce(0, torch.from_numpy(rowptr), None, dev(msgs))


# ------------------------------------------------------------------ GRU kernels
def gru_weights(rng, fi, u):
    lim = np.sqrt(6.0 / (fi + 3 * u))
    return (rng.uniform(-lim, lim, (fi, 3 * u)).astype(np.float32),
            rng.uniform(-lim, lim, (u, 3 * u)).astype(np.float32),
            rng.uniform(-0.1, 0.1, (2, 3 * u)).astype(np.float32))


@pytest.mark.parametrize("fi,u", [(32, 32), (64, 64), (16, 16), (16, 32), (64, 32), (32, 64), (32, 16)])
@pytest.mark.parametrize("n", [1, 129, 5000])
def test_gru_cell(fi, u, n):
    from ignnition_b200 import ops
    rng = np.random.RandomState(fi + u + n)
    K, R, b = gru_weights(rng, fi, u)
    x = rng.randn(n, fi).astype(np.float32)
    h = rng.randn(n, u).astype(np.float32)
    want = orc.gru_cell(x.astype(np.float64), h.astype(np.float64), K.astype(np.float64), R.astype(np.float64),
                        b.astype(np.float64))
    for tc in (True, False):          # tcgen05 3xTF32 (32 / 64 wide, n >= 128) and the fp32 twin
        got = ops.gru_cell(dev(x), dev(h), dev(K), dev(R), dev(b), tensor_cores=tc).cpu().numpy()
        assert rel_err(got, want) < RTOL, tc


@pytest.mark.parametrize("n,fi,u", [(4, 48, 48), (700, 20, 100), (5000, 128, 24)])
def test_gru_cell_generic_width(n, fi, u):
    """widths the fused kernels do not cover (hidden_state_dimension is free in the reference's schema) run as two
    Dense GEMMs + ign_gru_gates_fwd; the ordered WALK stays limited and says so"""
    from ignnition_b200 import ops
    rng = np.random.RandomState(n + u)
    x = rng.randn(n, fi).astype(np.float32)
    h = rng.randn(n, u).astype(np.float32)
    K, R, b = gru_weights(rng, fi, u)
    want = orc.gru_cell(x.astype(np.float64), h.astype(np.float64), K.astype(np.float64), R.astype(np.float64),
                        b.astype(np.float64))
    got = ops.gru_cell(dev(x), dev(h), dev(K), dev(R), dev(b)).cpu().numpy()
    assert rel_err(got, want) < RTOL
    # the ordered walk at that width: two sources, zero messages, empty destinations, per-step states
    lens = rng.randint(0, 6, n)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n_steps = int(lens.sum())
    if n_steps == 0:
        return
    rows = [max(2, n // 3), max(2, n // 5)]
    e_src = rng.randint(0, 2, n_steps)
    e_row = np.array([rng.randint(0, rows[k]) for k in e_src])
    steps = ((e_src.astype(np.int64) << 28) | e_row).astype(np.int32)
    zero = rng.rand(n_steps) < 0.1
    steps[zero] = -1
    states = [(rng.randn(r, fi) * 0.5).astype(np.float32) for r in rows]
    h_seq = torch.zeros(n_steps, u, device="cuda")
    got = ops.gru_seq(dev(rowptr, torch.int32), dev(steps, torch.int32), None, [dev(a) for a in states], dev(h), dev(K),
                      dev(R), dev(b), h_seq=h_seq).cpu().numpy()
    L = int(lens.max())
    padded = np.zeros((n, L, fi))
    d_of = np.repeat(np.arange(n), lens)
    t_of = np.arange(n_steps) - rowptr[d_of]
    msgs = np.zeros((n_steps, fi))
    for k in range(2):
        mk = (e_src == k) & ~zero
        msgs[mk] = states[k][e_row[mk]]
    padded[d_of, t_of] = msgs
    K64, R64, b64 = K.astype(np.float64), R.astype(np.float64), b.astype(np.float64)
    cell = lambda a_, h_: orc.gru_cell(a_, h_, K64, R64, b64)
    nz = lens > 0
    want = h.astype(np.float64).copy()
    want[nz] = orc.masked_rnn_last(cell, padded[nz], h[nz].astype(np.float64), lens[nz])
    assert rel_err(got, want) < RTOL
    assert np.array_equal(got[~nz], h[~nz])
    assert np.array_equal(h_seq.cpu().numpy()[rowptr[1:][nz] - 1], got[nz])


@pytest.mark.parametrize("fi,u", [(32, 32), (64, 64), (16, 32)])
def test_agg_gru_cell(fi, u):
    from ignnition_b200 import ops
    rng = np.random.RandomState(fi * u)
    n_dst, n_src = 2500, 900
    src, dst, seq = random_edges(rng, n_dst, n_src, 45)
    states = (rng.randn(n_src, fi) * 0.3).astype(np.float32)
    h = rng.randn(n_dst, u).astype(np.float32)
    K, R, b = gru_weights(rng, fi, u)
    r, c, _ = orc.csr_from_edges(src, dst, seq, n_dst)
    agg_out = torch.empty(n_dst, fi, device="cuda")
    got = ops.agg_gru_cell(dev(r, torch.int32), dev(c, torch.int32), dev(states), dev(h), dev(K), dev(R), dev(b),
                           agg_out=agg_out).cpu().numpy()
    agg = np.zeros((n_dst, fi), np.float64)
    np.add.at(agg, dst, states[src].astype(np.float64))
    want = orc.gru_cell(agg, h.astype(np.float64), K.astype(np.float64), R.astype(np.float64), b.astype(np.float64))
    assert rel_err(agg_out.cpu().numpy(), agg) < RTOL
    assert rel_err(got, want) < RTOL


@pytest.fixture
def fp32_kernels():
    """Run a test on the fp32 CUDA-core twins instead of the tcgen05 (3xTF32) kernels."""
    from ignnition_b200 import ops
    prev = ops.set_tensor_cores(False)
    yield
    ops.set_tensor_cores(prev)


def test_gru_seq_fp32_twin(fp32_kernels):
    test_gru_seq_vs_masked_rnn(32, 32, True)


@pytest.mark.parametrize("fi,u", [(32, 32), (64, 64)])
@pytest.mark.parametrize("use_order", [False, True])
def test_gru_seq_vs_masked_rnn(fi, u, use_order):
    """ordered aggregation: CSR walk == keras RNN over the dense right-padded tensor + mask.
    (32, 32) runs the tcgen05 3xTF32 kernel by default."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(fi + 7 * use_order)
    n_dst, n_src, max_len = 1500, 400, 7
    lens = rng.randint(0, max_len + 1, n_dst)      # includes empty destinations (state carried)
    dst = np.repeat(np.arange(n_dst), lens)
    seq = np.concatenate([np.arange(l) for l in lens])
    src = rng.randint(0, n_src, len(dst))
    states = (rng.randn(n_src, fi) * 0.5).astype(np.float32)
    h0 = rng.randn(n_dst, u).astype(np.float32)
    K, R, b = gru_weights(rng, fi, u)
    r, c, _ = orc.csr_from_edges(src, dst, seq, n_dst)
    rp, cc = dev(r, torch.int32), dev(c, torch.int32)
    order = ops.length_order(rp) if use_order else None
    h_seq = torch.zeros(len(dst), u, device="cuda")
    got = ops.gru_seq(rp, cc, order, [dev(states)], dev(h0), dev(K), dev(R), dev(b), h_seq=h_seq).cpu().numpy()
    padded = np.zeros((n_dst, max_len, fi), np.float64)
    padded[dst, seq] = states[src]
    K64, R64, b64 = K.astype(np.float64), R.astype(np.float64), b.astype(np.float64)
    cell = lambda a, h: orc.gru_cell(a, h, K64, R64, b64)
    nz = lens > 0
    want = h0.astype(np.float64).copy()
    want[nz] = orc.masked_rnn_last(cell, padded[nz], h0[nz].astype(np.float64), lens[nz])
    assert rel_err(got, want) < RTOL
    assert np.array_equal(got[~nz], h0[~nz])
    # saved per-step states: last step of every non-empty destination equals its new state
    hs = h_seq.cpu().numpy()
    assert np.array_equal(hs[r[1:][nz] - 1], got[nz])
    # a long-sequence case (link-like fan-in) through the same kernel
    lens2 = rng.randint(1, 60, 300)
    dst2 = np.repeat(np.arange(300), lens2)
    seq2 = np.concatenate([np.arange(l) for l in lens2])
    src2 = rng.randint(0, n_src, len(dst2))
    r2, c2, _ = orc.csr_from_edges(src2, dst2, seq2, 300)
    got2 = ops.gru_seq(dev(r2, torch.int32), dev(c2, torch.int32), None, [dev(states)], dev(h0[:300]), dev(K), dev(R),
                       dev(b)).cpu().numpy()
    pad2 = np.zeros((300, 59, fi), np.float64)
    pad2[dst2, seq2] = states[src2]
    want2 = orc.masked_rnn_last(cell, pad2, h0[:300].astype(np.float64), lens2)
    assert rel_err(got2, want2) < RTOL


@pytest.mark.parametrize("n_dst,max_len,n_srcs", [(1, 3, 1), (127, 1, 1), (129, 0, 1), (3000, 9, 1), (70000, 5, 2),
                                                  (1000, 33, 3)])
def test_gru_seq_proj(n_dst, max_len, n_srcs, monkeypatch):
    """ign_gru_seq_proj (input projection hoisted, h operand in tensor memory, three walkers per SM) == the masked
    RNN of the oracle (fp64) to 1e-5 and == ign_gru_seq to fp32 rounding: several source arrays, zero messages,
    empty destinations, ragged last tile, many tiles per walker, saved per-step states."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(n_dst + max_len)
    u = 32
    rows = [max(2, n_dst // 40 + 3 * k) for k in range(n_srcs)]
    lens = rng.randint(0, max_len + 1, n_dst)
    if n_dst > 5 and max_len > 0:
        lens[:5] = max_len
    n_steps = int(lens.sum())
    ent_src = rng.randint(0, n_srcs, n_steps)
    ent_row = np.array([rng.randint(0, rows[k]) for k in ent_src], dtype=np.int64) if n_steps else np.zeros(0, np.int64)
    steps = ((ent_src.astype(np.int64) << 28) | ent_row).astype(np.int32)
    zero = rng.rand(n_steps) < 0.05
    steps[zero] = -1
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    states = [(rng.randn(r, u) * 0.5).astype(np.float32) for r in rows]
    h0 = rng.randn(n_dst, u).astype(np.float32)
    K, R, b = gru_weights(rng, u, u)
    rp, st = dev(rowptr, torch.int32), dev(steps if n_steps else np.full(1, -1, np.int32), torch.int32)
    order = ops.length_order(rp)
    meta = ops.seq_meta(rp, st, order)
    srcs = [dev(x) for x in states]
    monkeypatch.setattr(ops, "gru_seq_proj_pays", lambda *a, **k: True)
    hs_p = torch.zeros(max(n_steps, 1), u, device="cuda")
    got = ops.gru_seq(rp, st, order, srcs, dev(h0), dev(K), dev(R), dev(b), h_seq=hs_p, meta=meta).cpu().numpy()
    monkeypatch.setattr(ops, "gru_seq_proj_pays", lambda *a, **k: False)
    hs_o = torch.zeros(max(n_steps, 1), u, device="cuda")
    old = ops.gru_seq(rp, st, order, srcs, dev(h0), dev(K), dev(R), dev(b), h_seq=hs_o, meta=meta).cpu().numpy()
    # oracle: the dense right-padded tensor + mask
    L = max(int(lens.max()) if n_dst else 0, 1)
    padded = np.zeros((n_dst, L, u), np.float64)
    d_of = np.repeat(np.arange(n_dst), lens)
    t_of = np.arange(n_steps) - rowptr[d_of]
    msgs = np.zeros((n_steps, u))
    for k in range(n_srcs):
        m = (ent_src == k) & ~zero
        msgs[m] = states[k][ent_row[m]]
    padded[d_of, t_of] = msgs
    K64, R64, b64 = K.astype(np.float64), R.astype(np.float64), b.astype(np.float64)
    cell = lambda a, h: orc.gru_cell(a, h, K64, R64, b64)
    nz = lens > 0
    want = h0.astype(np.float64).copy()
    if nz.any():
        want[nz] = orc.masked_rnn_last(cell, padded[nz], h0[nz].astype(np.float64), lens[nz])
    assert rel_err(got, want) < RTOL
    assert rel_err(got, old.astype(np.float64)) < RTOL
    assert np.array_equal(got[~nz], h0[~nz])
    if n_steps:
        assert np.array_equal(hs_p.cpu().numpy()[rowptr[1:][nz] - 1], got[nz])
        assert rel_err(hs_p.cpu().numpy(), hs_o.cpu().numpy().astype(np.float64)) < RTOL


@pytest.mark.parametrize("max_len", [1, 6, 16])
def test_gru_seq_step_synchronous(max_len):
    """step-synchronous launches == the sequence walk == masked RNN of the oracle (incl. empty destinations,
    zero messages and the saved per-step states)"""
    from ignnition_b200 import ops
    rng = np.random.RandomState(max_len)
    n_dst, n_src, u = 3000, 500, 32
    lens = rng.randint(0, max_len + 1, n_dst)
    lens[:5] = max_len
    dst = np.repeat(np.arange(n_dst), lens)
    seq = np.concatenate([np.arange(l) for l in lens])
    src = rng.randint(0, n_src, len(dst))
    states = (rng.randn(n_src, u) * 0.5).astype(np.float32)
    h0 = rng.randn(n_dst, u).astype(np.float32)
    K, R, b = gru_weights(rng, u, u)
    r, c, _ = orc.csr_from_edges(src, dst, seq, n_dst)
    c = c.copy()
    zero_slots = rng.rand(len(c)) < 0.05
    c[zero_slots] = -1                                  # IGN_STEP_ZERO entries: zero messages inside a sequence
    rp, cc = dev(r, torch.int32), dev(c, torch.int32)
    order = ops.length_order(rp)
    meta = ops.seq_meta(rp, cc, order)
    plan = ops.seq_step_plan(meta, cc, max_len)
    nt = plan[0].cpu().numpy()
    assert np.array_equal(nt, [(lens > t).sum() for t in range(max_len)])
    h_seq = torch.zeros(len(dst), u, device="cuda")
    got = ops.gru_seq_steps(plan, meta, [dev(states)], dev(h0), dev(K), dev(R), dev(b), max_len, h_seq=h_seq).cpu().numpy()
    h_seq2 = torch.zeros(len(dst), u, device="cuda")
    walk = ops.gru_seq(rp, cc, order, [dev(states)], dev(h0), dev(K), dev(R), dev(b), h_seq=h_seq2, meta=meta).cpu().numpy()
    padded = np.zeros((n_dst, max_len, u), np.float64)
    msgs = np.where(c[:, None] >= 0, states[np.maximum(c, 0)], 0.0)
    d_sorted = np.repeat(np.arange(n_dst), lens)
    padded[d_sorted, np.arange(len(c)) - r[d_sorted]] = msgs
    K64, R64, b64 = K.astype(np.float64), R.astype(np.float64), b.astype(np.float64)
    cell = lambda a, h: orc.gru_cell(a, h, K64, R64, b64)
    nz = lens > 0
    want = h0.astype(np.float64).copy()
    want[nz] = orc.masked_rnn_last(cell, padded[nz], h0[nz].astype(np.float64), lens[nz])
    assert rel_err(got, want) < RTOL and rel_err(walk, want) < RTOL
    assert np.array_equal(got[~nz], h0[~nz])
    assert np.array_equal(h_seq.cpu().numpy()[r[1:][nz] - 1], got[nz])
    assert rel_err(h_seq.cpu().numpy(), h_seq2.cpu().numpy().astype(np.float64)) < RTOL


def test_gru_seq_interleave_step_table(golden):
    """Q-size step 1: step table from ign_steps_build == reference interleave of the padded tensors."""
    from ignnition_b200 import ops
    from ignnition_b200.batching import AdjacencySpec, SequenceSpec, position_table
    g = golden("qsize_nsfnet")
    ref = g["reference_tensors"][0]
    adjs = [AdjacencySpec("adj_links_paths", "link", "path"), AdjacencySpec("adj_nodes_paths", "node", "path")]
    spec = SequenceSpec("k", "path", adjs, True)
    ps, pc = position_table(ref, spec)
    n = ref["num_path"]
    rps, cols, np_csr = [], [], []
    for a in adjs:
        r, c, _ = orc.csr_from_edges(ref["src_" + a.name], ref["dst_" + a.name], ref[a.seq_key], n)
        np_csr.append((r, c))
        rps.append(dev(r, torch.int32)); cols.append(dev(c, torch.int32))
    total = sum(len(c) for _, c in np_csr)
    srp, steps = ops.steps_build(rps, cols, None, dev(np.array([0, len(ps)], np.int32)), dev(ps), dev(pc), n, total)
    srp, steps = srp.cpu().numpy(), steps.cpu().numpy()
    # reference: concat padded blocks, scatter columns to `indices`, first final_len columns are the sequence
    idx = np.concatenate([ref["indices_link_to_path"], ref["indices_node_to_path"]])
    lens = [np.diff(r) for r, _ in np_csr]
    maxl = [int(l.max()) for l in lens]
    for d in range(n):
        cols_d = []
        for k, (r, c) in enumerate(np_csr):
            row = [(k << 28) | int(x) for x in c[r[d]:r[d + 1]]] + [-1] * (maxl[k] - lens[k][d])
            cols_d += row
        seq_d = [-1] * len(idx)
        for cpos, p in enumerate(idx):
            seq_d[p] = cols_d[cpos]
        fl = lens[0][d] + lens[1][d]
        assert srp[d + 1] - srp[d] == fl
        assert steps[srp[d]:srp[d + 1]].tolist() == seq_d[:fl]


# ------------------------------------------------------------------ dense layers
@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("m,k,n,act", [(1, 32, 256, "selu"), (1000, 32, 256, "selu"), (777, 256, 256, "relu"),
                                       (5000, 256, 1, None), (130, 65, 20, "tanh"), (300, 7, 3, "sigmoid"),
                                       (128, 32, 32, None), (40000, 64, 96, "selu"), (2049, 256, 128, "tanh"),
                                       # M >= 4096 and N % 64 == 0: the warp-specialised pipeline (dense_pipe_tc_kernel)
                                       (5003, 256, 256, "selu"), (4097, 32, 64, "tanh"), (9000, 96, 192, "relu"),
                                       (4500, 64, 128, None), (20000, 32, 256, "selu"), (7001, 256, 32, None), (4100, 64, 32, "selu")])
def test_dense(m, k, n, act, tensor_cores):
    """tensor_cores=True: 3xTF32 on tcgen05 where the shape is built (K % 32 == 0, N % 32 == 0, M >= 128),
    else the fp32 CUDA-core kernel; both must meet the fp32 parity bar."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(m + k + n)
    x = rng.randn(m, k).astype(np.float32)
    w = (rng.randn(k, n) / np.sqrt(k)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, n).astype(np.float32)
    pre = torch.empty(m, n, device="cuda")
    got = ops.dense(dev(x), dev(w), dev(b), ops.ACTIVATIONS[act], pre_act=pre,
                    tensor_cores=tensor_cores).cpu().numpy()
    z = x.astype(np.float64) @ w.astype(np.float64) + b
    assert rel_err(pre.cpu().numpy(), z) < RTOL
    assert rel_err(got, orc.activation(act, z)) < RTOL


def test_dense_head_fused():
    """readout tail: out = selu(x W + b) . w3 + b3 in one kernel == two dense layers"""
    from ignnition_b200 import ops
    rng = np.random.RandomState(11)
    m, k, n = 3000, 256, 256
    x = rng.randn(m, k).astype(np.float32)
    w = (rng.randn(k, n) / np.sqrt(k)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, n).astype(np.float32)
    w3 = (rng.randn(n) / np.sqrt(n)).astype(np.float32)
    b3 = np.array([0.37], np.float32)
    assert ops.dense_head_supported(m, k, n)
    got = ops.dense_head(dev(x), dev(w), dev(b), ops.ACTIVATIONS["selu"], dev(w3), dev(b3)).cpu().numpy()
    want = orc.activation("selu", x.astype(np.float64) @ w.astype(np.float64) + b) @ w3.astype(np.float64) + b3[0]
    assert got.shape == (m, 1) and rel_err(got.reshape(-1), want) < RTOL


@pytest.mark.parametrize("m,k1,n1,n2", [(3000, 32, 256, 256), (128, 64, 96, 32), (70001, 32, 256, 256)])
def test_mlp_head_fused(m, k1, n1, n2):
    """whole readout stack in one kernel == three dense layers of the oracle"""
    from ignnition_b200 import ops
    rng = np.random.RandomState(m % 97)
    x = rng.randn(m, k1).astype(np.float32)
    w1 = (rng.randn(k1, n1) / np.sqrt(k1)).astype(np.float32); b1 = rng.uniform(-0.1, 0.1, n1).astype(np.float32)
    w2 = (rng.randn(n1, n2) / np.sqrt(n1)).astype(np.float32); b2 = rng.uniform(-0.1, 0.1, n2).astype(np.float32)
    w3 = (rng.randn(n2) / np.sqrt(n2)).astype(np.float32); b3 = np.array([-0.21], np.float32)
    assert ops.mlp_head_supported(m, k1, n1, n2)
    got = ops.mlp_head(dev(x), dev(w1), dev(b1), ops.ACTIVATIONS["selu"], dev(w2), dev(b2), ops.ACTIVATIONS["selu"],
                       dev(w3), dev(b3)).cpu().numpy()
    h1 = orc.activation("selu", x.astype(np.float64) @ w1.astype(np.float64) + b1)
    h2 = orc.activation("selu", h1 @ w2.astype(np.float64) + b2)
    want = h2 @ w3.astype(np.float64) + b3[0]
    assert got.shape == (m, 1) and rel_err(got.reshape(-1), want) < RTOL


def test_attention_aggregate_vs_padded_softmax():
    """ign_attention_aggregate against the reference formulation on the padded tensor: softmax over the
    DESTINATIONS of one sample per padded column, zero pads included (auxilary_classes.py:318-338)."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(4)
    F = 32
    samples = [(7, 5), (1, 3), (40, 25)]                 # (destinations, sources) per sample; block-diagonal batch
    dsts, srcs, doff, soff = [], [], [0], [0]
    for nd, ns in samples:
        for d in range(nd):
            k = rng.randint(0, 5)                        # some destinations receive nothing
            dsts += [doff[-1] + d] * k
            srcs += list(soff[-1] + rng.randint(0, ns, k))
        doff.append(doff[-1] + nd); soff.append(soff[-1] + ns)
    dst = np.asarray(dsts, np.int32); src = np.asarray(srcs, np.int32)
    n_dst, n_src = doff[-1], soff[-1]
    rows = rng.randn(n_src, F).astype(np.float32)
    s_src = rng.randn(n_src).astype(np.float32); s_dst = rng.randn(n_dst).astype(np.float32)
    rowptr, col, _, _ = ops.csr_build(dev(dst), dev(src), None, n_dst)
    deg = np.bincount(dst, minlength=n_dst)
    got = ops.attention_aggregate(rowptr, col, dev(rows), dev(s_src), dev(s_dst), dev(np.asarray(doff, np.int32)),
                                  int(deg.max())).cpu().numpy()
    want = np.zeros((n_dst, F))
    for k in range(len(samples)):
        d0, d1 = doff[k], doff[k + 1]
        L = max(int(deg[d0:d1].max()), 1)
        aux = np.zeros((d1 - d0, L)); msg = np.zeros((d1 - d0, L, F)); pos = np.zeros(d1 - d0, int)
        for e in range(len(dst)):
            if d0 <= dst[e] < d1:
                a = float(s_src[src[e]]) + float(s_dst[dst[e]])
                aux[dst[e] - d0, pos[dst[e] - d0]] = a if a > 0 else 0.2 * a
                msg[dst[e] - d0, pos[dst[e] - d0]] = rows[src[e]]
                pos[dst[e] - d0] += 1
        ex = np.exp(aux - aux.max(axis=0, keepdims=True))
        coef = ex / ex.sum(axis=0, keepdims=True)        # softmax over axis 0
        valid = np.arange(L)[None, :] < deg[d0:d1, None]
        want[d0:d1] = ((coef * valid)[:, :, None] * msg).sum(axis=1)
    assert rel_err(got, want) < RTOL


def test_partner_index_conv_finish_mul():
    from ignnition_b200 import ops
    rng = np.random.RandomState(6)
    rp0 = np.asarray([0, 2, 2, 5, 6], np.int32); rp1 = np.asarray([0, 1, 4, 6, 6], np.int32)
    idx1 = np.asarray([10, 11, 12, 13, 14, 15], np.int32)
    got = ops.partner_index(dev(rp0), dev(rp1), dev(idx1), 6).cpu().numpy()
    assert np.array_equal(got, [10, -1, 14, 15, -1, -1])     # bit-exact index work
    n, F = 50, 12
    nsum = rng.randn(n, F).astype(np.float32); me = rng.randn(n, F).astype(np.float32)
    deg = rng.randint(1, 6, n); rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    got = ops.conv_finish(dev(nsum), dev(me), dev(rowptr), ops.ACTIVATIONS["relu"]).cpu().numpy()
    assert rel_err(got, np.maximum((nsum.astype(np.float64) + me) / deg[:, None], 0)) < RTOL
    a = rng.randn(33, 7).astype(np.float32); b = rng.randn(33, 7).astype(np.float32)
    assert np.array_equal(ops.mul(dev(a), dev(b)).cpu().numpy(), a * b)
    # a negative index gathers a zero row (padding of the concat-axis-2 layout)
    out = ops.gather_concat([dev(a)], [dev(np.asarray([0, -1, 32], np.int32))], 3).cpu().numpy()
    assert np.array_equal(out, np.stack([a[0], np.zeros(7, np.float32), a[32]]))


def test_csr_build_presorted_and_unsorted_agree():
    """the sorted-input fast path (no radix passes) and the sort give the same CSR, bit for bit"""
    from ignnition_b200 import ops
    rng = np.random.RandomState(9)
    n_dst, E = 5000, 60000
    dst_sorted = np.sort(rng.randint(0, n_dst, E)).astype(np.int32)
    src = rng.randint(0, 777, E).astype(np.int32)
    rp, col, perm, _ = ops.csr_build(dev(dst_sorted), dev(src), None, n_dst, want_perm=True)
    want_rp, want_col, want_perm = orc.stable_sort_csr(src, dst_sorted, n_dst)
    assert np.array_equal(rp.cpu().numpy(), want_rp) and np.array_equal(col.cpu().numpy(), want_col)
    assert np.array_equal(perm.cpu().numpy(), np.arange(E))
    shuffle = rng.permutation(E)
    rp2, col2, perm2, _ = ops.csr_build(dev(dst_sorted[shuffle]), dev(src[shuffle]), None, n_dst, want_perm=True)
    w_rp, w_col, w_perm = orc.stable_sort_csr(src[shuffle], dst_sorted[shuffle], n_dst)
    assert np.array_equal(rp2.cpu().numpy(), w_rp) and np.array_equal(col2.cpu().numpy(), w_col)
    assert np.array_equal(perm2.cpu().numpy(), w_perm)


def test_init_state_and_gather_concat():
    from ignnition_b200 import ops
    rng = np.random.RandomState(0)
    a = rng.randn(100, 1).astype(np.float32)
    b = rng.randn(100, 3).astype(np.float32)
    st = ops.init_state([dev(a), dev(b)], [1, 3], 100, 32).cpu().numpy()
    assert np.array_equal(st[:, :1], a) and np.array_equal(st[:, 1:4], b) and np.all(st[:, 4:] == 0)
    idx = rng.randint(0, 100, 250).astype(np.int32)
    p = rng.randn(250, 2).astype(np.float32)
    out = ops.gather_concat([dev(st), dev(p)], [dev(idx), None], 250).cpu().numpy()
    assert np.array_equal(out[:, :32], st[idx]) and np.array_equal(out[:, 32:], p)


@pytest.mark.parametrize("m,k,n,act", [(5003, 32, 256, "selu"), (9001, 256, 256, "selu"), (4500, 64, 128, "tanh"),
                                       (4096, 128, 64, "relu"), (6000, 256, 32, None), (777, 32, 256, "selu"),
                                       (5000, 256, 1, None)])
def test_dense_bwd(m, k, n, act):
    """tf.gradients through Dense (generate_model.py:791): dX = dZ W^T, dW += X^T dZ, db += colsum(dZ),
    dZ = dY * act'(pre).  m >= 4096 with 32-multiples runs dW on tcgen05 (MN-major operands, dw_tc.cu),
    the rest on the fp32 kernels; both against fp64."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(m + k + n)
    x = rng.randn(m, k).astype(np.float32)
    w = (rng.randn(k, n) / np.sqrt(k)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, n).astype(np.float32)
    dy = rng.randn(m, n).astype(np.float32)
    pre = x.astype(np.float64) @ w.astype(np.float64) + b
    a = ops.ACTIVATIONS[act]
    if act == "selu":
        s, al = 1.0507009873554805, 1.6732632423543772
        d = np.where(pre > 0, s, s * al * np.exp(pre))
    elif act == "tanh":
        d = 1 - np.tanh(pre) ** 2
    elif act == "relu":
        d = (pre > 0).astype(np.float64)
    else:
        d = np.ones_like(pre)
    dz = dy.astype(np.float64) * d
    want_dx = dz @ w.astype(np.float64).T
    want_dw = x.astype(np.float64).T @ dz
    want_db = dz.sum(0)
    dx = torch.empty(m, k, dtype=torch.float32, device="cuda")
    dw = torch.zeros(k, n, dtype=torch.float32, device="cuda")
    db = torch.zeros(n, dtype=torch.float32, device="cuda")
    ops.dense_bwd(dev(x), dev(w), a, dev(pre.astype(np.float32)), dev(dy), dx, dw, db)
    assert rel_err(dx.cpu().numpy(), want_dx) < RTOL
    assert rel_err(dw.cpu().numpy(), want_dw) < RTOL
    assert rel_err(db.cpu().numpy(), want_db) < RTOL
    # gradients accumulate: a second call doubles dW and db; this one takes act' from the layer's OUTPUT
    # (IGN_ACT_FROM_OUTPUT), which is what the train step saves
    y = orc.activation(act, pre).astype(np.float32)
    ops.dense_bwd(dev(x), dev(w), a | ops.ACT_FROM_OUTPUT, dev(y), dev(dy), dx, dw, db)
    assert rel_err(dx.cpu().numpy(), want_dx) < 2 * RTOL
    assert rel_err(dw.cpu().numpy(), 2 * want_dw) < 2 * RTOL


@pytest.mark.parametrize("max_len,n_dst", [(1, 700), (6, 3000), (16, 1000), (3, 90000)])
def test_gru_seq_bwd_step_synchronous(max_len, n_dst):
    """BPTT as step-synchronous tcgen05 launches (ign_gru_seq_bwd_steps) == torch.autograd in fp64 through the
    oracle's GRU recurrence, and == the fp32 tile-walk kernel (ign_gru_seq_bwd): message gradients per step,
    initial-state gradients (incl. destinations without steps), kernel / recurrent kernel / bias gradients;
    two sources and zero-message entries in the step table."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(max_len + n_dst)
    n_src, u = 400, 32
    lens = rng.randint(0, max_len + 1, n_dst)
    lens[:5] = max_len
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n_steps = int(rowptr[-1])
    which = rng.randint(0, 2, n_steps)
    rows = rng.randint(0, n_src, n_steps)
    steps = ((which << ops.STEP_SRC_SHIFT) | rows).astype(np.int32)
    steps[rng.rand(n_steps) < 0.05] = -1
    s0 = (rng.randn(n_src, u) * 0.5).astype(np.float32)
    s1 = (rng.randn(n_src, u) * 0.5).astype(np.float32)
    h0 = (rng.randn(n_dst, u) * 0.7).astype(np.float32)
    d_out = rng.randn(n_dst, u).astype(np.float32)
    K, R, b = gru_weights(rng, u, u)

    # fp64 autograd through h <- GRU(x_t, h) (Keras GRUCell v2, reset_after=True)
    tK, tR, tb = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (K, R, b)]
    th0 = torch.tensor(h0, dtype=torch.float64, requires_grad=True)
    src_all = torch.tensor(np.concatenate([s0, s1, np.zeros((1, u), np.float32)]), dtype=torch.float64)
    idx = np.where(steps >= 0, (steps >> ops.STEP_SRC_SHIFT) * n_src + (steps & ((1 << ops.STEP_SRC_SHIFT) - 1)), 2 * n_src)
    x_all = src_all[torch.from_numpy(idx)].clone().requires_grad_(True)
    h = th0
    lens_t = torch.from_numpy(lens)
    rp_t = torch.from_numpy(rowptr[:-1].astype(np.int64))
    for t in range(max_len):
        act = lens_t > t
        pos = (rp_t + t)[act]
        x = x_all[pos]
        hp = h[act]
        mx = x @ tK + tb[0]
        mh = hp @ tR + tb[1]
        z = torch.sigmoid(mx[:, :u] + mh[:, :u])
        r = torch.sigmoid(mx[:, u:2 * u] + mh[:, u:2 * u])
        hh = torch.tanh(mx[:, 2 * u:] + r * mh[:, 2 * u:])
        hn = z * hp + (1 - z) * hh
        h = h.clone()
        h[act] = hn
    (h * torch.tensor(d_out, dtype=torch.float64)).sum().backward()

    rp, st = dev(rowptr), dev(steps)
    order = ops.length_order(rp)
    meta = ops.seq_meta(rp, st, order)
    plan = ops.seq_step_plan(meta, st, max_len)
    srcs = [dev(s0), dev(s1)]
    h_seq = torch.zeros(max(n_steps, 1), u, device="cuda")
    ops.gru_seq(rp, st, order, srcs, dev(h0), dev(K), dev(R), dev(b), h_seq=h_seq, meta=meta)

    def run(tc):
        d_steps = torch.zeros(max(n_steps, 1), u, device="cuda")
        dh0 = torch.zeros(n_dst, u, device="cuda")
        dk, dr, db = torch.zeros(u, 3 * u, device="cuda"), torch.zeros(u, 3 * u, device="cuda"), torch.zeros(2, 3 * u, device="cuda")
        if tc:
            ops.gru_seq_bwd_steps(plan, meta, max_len, srcs, dev(h0), h_seq, dev(K), dev(R), dev(b), dev(d_out),
                                  d_steps, dh0, dk, dr, db)
        else:
            ops.gru_seq_bwd(rp, st, order, srcs, dev(h0), h_seq, dev(K), dev(R), dev(b), dev(d_out), d_steps, dh0,
                            dk, dr, db)
        return [a.cpu().numpy() for a in (d_steps[:n_steps], dh0, dk, dr, db)]

    want = [x_all.grad.numpy(), th0.grad.numpy(), tK.grad.numpy(), tR.grad.numpy(), tb.grad.numpy()]
    for tc in (False, True):
        got = run(tc)
        for name, g, w in zip(("d_steps", "dh0", "dK", "dR", "db"), got, want):
            assert rel_err(g, w) < 2e-5, (tc, name, rel_err(g, w))


@pytest.mark.parametrize("m,k,act", [(5000, 256, "selu"), (1500, 128, "tanh"), (3000, 512, "relu"), (2048, 256, None)])
def test_dense_head_bwd_chain(m, k, act):
    """backward of a linear k -> 1 head chained with the layer below (ign_dense_head_bwd_chain): dw += x^T dz,
    dz_prev = (dz w^T) * act'(x) with act' taken from x = the lower layer's output, db_prev += colsum(dz_prev); vs fp64"""
    from ignnition_b200 import ops
    rng = np.random.RandomState(m + k)
    pre = rng.randn(m, k)
    x = orc.activation(act, pre).astype(np.float32)             # output of the layer below
    w = (rng.randn(k) / np.sqrt(k)).astype(np.float32)
    dz = rng.randn(m).astype(np.float32)
    x64 = x.astype(np.float64)
    if act == "selu":
        s_, al = 1.0507009873554805, 1.6732632423543772
        d = np.where(x64 > 0, s_, x64 + s_ * al)
    elif act == "tanh":
        d = 1 - x64 ** 2
    elif act == "relu":
        d = (x64 > 0).astype(np.float64)
    else:
        d = np.ones_like(x64)
    want_prev = dz.astype(np.float64)[:, None] * w.astype(np.float64)[None, :] * d
    want_dw = x64.T @ dz.astype(np.float64)
    dz_prev = torch.empty(m, k, device="cuda")
    dw = torch.zeros(k, device="cuda")
    db = torch.zeros(k, device="cuda")
    assert ops.dense_head_bwd_chain_supported(m, k, 1)
    ops.dense_head_bwd_chain(dev(x), dev(w), dev(dz), ops.ACTIVATIONS[act], dz_prev, dw, db)
    assert rel_err(dz_prev.cpu().numpy(), want_prev) < RTOL
    assert rel_err(dw.cpu().numpy(), want_dw) < RTOL
    assert rel_err(db.cpu().numpy(), want_prev.sum(0)) < RTOL


# ------------------------------------------------------------------ fused gather + aggregation + GRU + TMA stores
def _skewed_edges(rng, n_dst, n_src, max_len, hub=0):
    src, dst, seq = random_edges(rng, n_dst, n_src, max_len)
    if hub and n_dst > 3:          # one destination with a long list (edge-balanced row ranges must cope)
        extra = rng.randint(0, n_src, hub)
        d = 3
        base = int((dst == d).sum())
        src = np.concatenate([src, extra])
        dst = np.concatenate([dst, np.full(hub, d)])
        seq = np.concatenate([seq, base + np.arange(hub)])
    return src, dst, seq


@pytest.mark.parametrize("u", [32, 64])
@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("n_dst,max_len,hub", [(1, 3, 0), (127, 5, 0), (128, 0, 0), (129, 9, 0), (1000, 40, 700),
                                              (20000, 12, 0)])
def test_agg_gru_cell_tc(u, op, n_dst, max_len, hub):
    """ign_agg_gru_cell_tc == ign_segment_reduce + ign_gru_cell (same sum order, same gate GEMMs: bit for bit)
    and == the fp64 oracle to 1e-5; two output arrays written at a row offset, rows past the range untouched."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(u + op + n_dst)
    n_src = max(3, n_dst // 2)
    src, dst, seq = _skewed_edges(rng, n_dst, n_src, max_len, hub)
    states = (rng.randn(n_src, u) * 0.3).astype(np.float32)
    h = rng.randn(n_dst, u).astype(np.float32)
    K, R, b = gru_weights(rng, u, u)
    r, c, _ = orc.csr_from_edges(src, dst, seq, n_dst)
    rp, cl, st, hd = dev(r, torch.int32), dev(c, torch.int32), dev(states), dev(h)
    row0, pad = 77, 200
    outs = [torch.full((row0 + n_dst + pad, u), 7.5, device="cuda") for _ in range(2)]
    agg_out = torch.empty(n_dst, u, device="cuda")
    ops.agg_gru_cell_tc(op, rp, cl, st, hd, dev(K), dev(R), dev(b), outs, out_row0=row0, agg_out=agg_out)
    agg_ref = ops.segment_reduce(op, rp, cl, st)
    assert torch.equal(agg_out, agg_ref)
    got = outs[0][row0:row0 + n_dst].cpu().numpy()
    assert torch.equal(outs[0], outs[1])
    assert float(outs[0][:row0].min()) == 7.5 and float(outs[0][row0 + n_dst:].max()) == 7.5
    want = orc.gru_cell(agg_ref.cpu().numpy().astype(np.float64), h.astype(np.float64), K.astype(np.float64),
                        R.astype(np.float64), b.astype(np.float64))
    assert rel_err(got, want) < RTOL
    if n_dst >= 128:   # the unfused tensor-core pair computes the same products in the same order
        pair = ops.gru_cell(agg_ref, hd, dev(K), dev(R), dev(b)).cpu().numpy()
        assert np.array_equal(got, pair)


def test_agg_gru_cell_tc_errors():
    from ignnition_b200 import ops
    z = torch.zeros(4, 48, device="cuda")
    rp = torch.zeros(5, dtype=torch.int32, device="cuda")
    cl = torch.zeros(1, dtype=torch.int32, device="cuda")
    w = torch.zeros(48, 144, device="cuda")
    with pytest.raises(RuntimeError, match="IGNNITION"):
        ops.agg_gru_cell_tc(0, rp, cl, z, z, w, w, torch.zeros(2, 144, device="cuda"), [torch.empty_like(z)])


def test_csr_rank_gaps_are_zero_rows():
    """a seq with gaps / duplicates leaves slots no edge claims; they must read as zero rows, not as
    uninitialised indices, and the status word must report them."""
    from ignnition_b200 import ops
    dst = np.array([0, 0, 0, 1, 1], np.int32)
    src = np.array([2, 1, 0, 3, 3], np.int32)
    seq = np.array([0, 0, 2, 0, 1], np.int32)          # duplicate 0, gap at 1 in destination 0
    rowptr, col, perm, status = ops.csr_build(dev(dst), dev(src), dev(seq), 2, ops.CSR_RANK, want_perm=True,
                                              want_status=True)
    c = col.cpu().numpy()
    assert c[1] == -1 and status.cpu().numpy()[0] > 0
    states = np.arange(4 * 4, dtype=np.float32).reshape(4, 4)
    got = ops.segment_reduce(0, rowptr, col, dev(states)).cpu().numpy()
    assert np.array_equal(got[0], states[c[0]] + states[0]) and np.array_equal(got[1], 2 * states[3])
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.index_range_check(dev(np.array([0, 5, -1, 3], np.int32)), 4, bad)
    assert int(bad.item()) == 2


@pytest.mark.parametrize("rows,widths,units,act", [(4096, [32], 32, "linear"), (4133, [32, 64], 64, "selu"),
                                                   (5000, [64, 64, 32], 128, "tanh"), (9000, [32, 32], 32, "relu"),
                                                   (4500, [64, 32, 32, 128], 256, "sigmoid")])
def test_gather_dense(rows, widths, units, act):
    """ign_gather_dense (gather + concat inside the A-operand loaders of the tcgen05 GEMM, small-M and pipelined
    kernels) == ign_gather_concat + ign_dense bit for bit (same GEMM on the same operand values) and == numpy fp64 to
    1e-5; identity index, negative indices (zero rows), ragged last tile."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(rows + units)
    parts, idx, parts_np, idx_np = [], [], [], []
    for k, wd in enumerate(widths):
        if k == 1:                      # a per-edge array read in place (identity index)
            a = rng.randn(rows, wd).astype(np.float32)
            ix = None
        else:
            n_src = max(3, rows // 7 + k)
            a = rng.randn(n_src, wd).astype(np.float32)
            ix = rng.randint(0, n_src, rows).astype(np.int32)
            ix[rng.rand(rows) < 0.03] = -1
        parts_np.append(a); idx_np.append(ix)
        parts.append(dev(a)); idx.append(None if ix is None else dev(ix, torch.int32))
    K = sum(widths)
    W = (rng.randn(K, units) / np.sqrt(K)).astype(np.float32)
    b = rng.randn(units).astype(np.float32) * 0.1
    a_id = ops.ACTIVATIONS[act]
    assert ops.gather_dense_supported(widths, units, rows)
    got = ops.gather_dense(parts, idx, rows, dev(W), dev(b), a_id).cpu().numpy()
    x = ops.gather_concat(parts, idx, rows)
    two = ops.dense(x, dev(W), dev(b), a_id).cpu().numpy()
    assert np.array_equal(got, two)
    x64 = np.concatenate([(a[np.maximum(ix, 0)] * (ix >= 0)[:, None]) if ix is not None else a
                          for a, ix in zip(parts_np, idx_np)], axis=1).astype(np.float64)
    assert np.array_equal(x.cpu().numpy(), x64.astype(np.float32))
    want = orc.ACT[act](x64 @ W.astype(np.float64) + b.astype(np.float64)) if hasattr(orc, "ACT") else None
    if want is None:
        z = x64 @ W.astype(np.float64) + b.astype(np.float64)
        want = {"linear": z, "relu": np.maximum(z, 0), "tanh": np.tanh(z), "sigmoid": 1 / (1 + np.exp(-z)),
                "selu": 1.0507009873554805 * np.where(z > 0, z, 1.6732632423543772 * (np.exp(z) - 1))}[act]
    assert rel_err(got, want) < RTOL
    assert not ops.gather_dense_supported([24, 32], units, rows) and not ops.gather_dense_supported(widths, units, 100)


def test_new_kernels_write_only_their_outputs():
    """compute-sanitizer is closed on the GPU pool (profiles/r2_sanitizer_unavailable.log), so the round-2 kernels are
    checked the way the pool suggests: outputs carved out of larger buffers filled with a sentinel, ragged sizes, and
    every byte outside the output compared afterwards (ign_gru_seq_proj with per-step states, ign_gather_dense on both
    GEMM kernels, ign_segment_max_bwd, ign_segment_broadcast)."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(9)
    u, pad, S = 32, 300, 1234.5

    def guarded(rows, width):
        buf = torch.full((rows + 2 * pad, width), S, device="cuda")
        return buf, buf[pad:pad + rows]

    def clean(buf, rows):
        return bool((buf[:pad] == S).all()) and bool((buf[pad + rows:] == S).all())

    for n_dst in (1, 129, 40000):
        lens = rng.randint(0, 5, n_dst)
        lens[0] = 4
        n_steps = int(lens.sum())
        rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        n_src = max(2, n_dst // 30)
        steps = rng.randint(0, n_src, n_steps).astype(np.int32)
        rp, st = dev(rowptr, torch.int32), dev(steps, torch.int32)
        order = ops.length_order(rp)
        meta = ops.seq_meta(rp, st, order)
        K, R, b = gru_weights(rng, u, u)
        ob, out = guarded(n_dst, u)
        hb, hs = guarded(n_steps, u)
        orig = ops.gru_seq_proj_pays
        ops.gru_seq_proj_pays = lambda *a, **k: True
        try:
            ops.gru_seq(rp, st, order, [dev((rng.randn(n_src, u) * 0.5).astype(np.float32))],
                        dev(rng.randn(n_dst, u).astype(np.float32)), dev(K), dev(R), dev(b), out=out, h_seq=hs, meta=meta)
        finally:
            ops.gru_seq_proj_pays = orig
        torch.cuda.synchronize()
        assert clean(ob, n_dst) and clean(hb, n_steps)
        assert bool(torch.isfinite(out).all()) and not bool((out == S).any())
    for rows in (129, 4097):
        parts = [dev(rng.randn(50, 32).astype(np.float32)), dev(rng.randn(rows, 32).astype(np.float32))]
        idx = [dev(rng.randint(0, 50, rows).astype(np.int32), torch.int32), None]
        yb, y = guarded(rows, 64)
        ops.gather_dense(parts, idx, rows, dev(rng.randn(64, 64).astype(np.float32)), dev(np.zeros(64, np.float32)), 0, out=y)
        torch.cuda.synchronize()
        assert clean(yb, rows) and not bool((y == S).any())


def test_csr_build_small_one_launch_bit_exact():
    """ign_csr_build_small (one CTA per adjacency, one launch for the whole graph) against the oracle's CSR: an
    adjacency placed by seq, one grouped by destination without seq, two in random order without seq (short rows: atomic placement + per-row sort of the edge ids;
    a row of ~800 edges: one warp walking the list), one without edges, and a seq with a gap (the unclaimed slot stays -1)."""
    from ignnition_b200 import ops
    rng = np.random.RandomState(17)
    src_a, dst_a, seq_a = random_edges(rng, 700, 300, 9)
    E = 5000
    dst_b = np.sort(rng.randint(0, 3000, E))
    src_b = rng.randint(0, 500, E)
    dst_c = rng.randint(0, 9000, E + 77)
    src_c = rng.randint(0, 500, E + 77)
    empty = np.zeros(0, np.int32)
    dst_e, seq_e, src_e = np.array([0, 0, 2]), np.array([0, 2, 0]), np.array([5, 6, 7])   # destination 0: slot 1 unclaimed
    i32 = lambda a: dev(np.asarray(a), torch.int32)
    dst_d = np.where(rng.rand(2000) < 0.4, 7, rng.randint(0, 50, 2000))          # a row of ~800 edges, any order
    src_d = rng.randint(0, 500, 2000)
    specs = [(i32(dst_a), i32(src_a), i32(seq_a), 700, True), (i32(dst_b), i32(src_b), None, 3000, True),
             (i32(dst_c), i32(src_c), None, 9000, True), (i32(empty), i32(empty), None, 12, False),
             (i32(dst_e), i32(src_e), i32(seq_e), 3, True), (i32(dst_d), i32(src_d), None, 50, True),
             (i32(dst_c), i32(src_c), None, 9000, False)]
    l0 = ops._lib.load().ign_launch_count()
    out = ops.csr_build_small(specs)
    assert ops._lib.load().ign_launch_count() - l0 == 1
    r, c, p = orc.csr_from_edges(src_a, dst_a, seq_a, 700)
    assert all(np.array_equal(x.cpu().numpy(), y) for x, y in zip(out[0], (r, c, p)))
    for k, (s_, d_, n_) in ((1, (src_b, dst_b, 3000)), (2, (src_c, dst_c, 9000)), (5, (src_d, dst_d, 50))):
        r, c, p = orc.stable_sort_csr(s_, d_, n_)
        assert all(np.array_equal(x.cpu().numpy(), y) for x, y in zip(out[k], (r, c, p))), k
    r, c, _ = orc.stable_sort_csr(src_c, dst_c, 9000)          # no perm asked for: the ids are sorted inside col
    assert np.array_equal(out[6][0].cpu().numpy(), r) and np.array_equal(out[6][1].cpu().numpy(), c) and out[6][2] is None
    assert np.array_equal(out[3][0].cpu().numpy(), np.zeros(13)) and out[3][2] is None
    assert np.array_equal(out[4][0].cpu().numpy(), [0, 2, 2, 3])
    assert np.array_equal(out[4][1].cpu().numpy(), [5, -1, 7]) and np.array_equal(out[4][2].cpu().numpy(), [0, -1, 2])
