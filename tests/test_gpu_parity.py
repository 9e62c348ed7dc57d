"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit at every ciphertext tap, and
against the reference's cleartext semantics after decryption.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

from conftest import (load_lbf_index, load_ref_mapped, out_hash, read_golden_blif, read_golden_lbf, selfcheck_inputs,
                      unpack_outputs)
from oracle import cleartext
from oracle.tfhe_ref import RefTFHE, GOLDILOCKS_P
from tfhe_fbs_map_b200 import levelize, params
from tfhe_fbs_map_b200.formats import read_lbf

pytestmark = pytest.mark.gpu
TOYS = ["toy1", "toy2", "toy3", "toy4", "toy5", "toy6", "toy2u", "toy3u", "toy5u", "toy7u", "toy3v", "toy5v"]     # *u / *v: two / three key bits per blind-rotation step (toy7u: odd n)
SEED = 777


@pytest.fixture(scope="module")
def ctxs():
    from tfhe_fbs_map_b200.backend import B200Backend
    cache = {}

    def get(name):
        if name not in cache:
            ps = params.get(name)
            cache[name] = (B200Backend(name, device=0, seed=SEED), RefTFHE(ps, seed=SEED))
        return cache[name]
    yield get
    for be, _ in cache.values():
        be.close()


def tables_for(p, rng):
    low = [int(x) for x in rng.integers(0, 2, p)]
    tabs = [(low, 1), (low + [1 - x for x in low], 1), ([0] + low[1:] + [0], 0), ([1] + low[1:] + [1], 2)]
    if p >= 3:
        tabs.append(([0, 2, 1], 1))
    return tabs


@pytest.mark.parametrize("name", TOYS)
def test_ntt_matches_oracle(ctxs, name):
    be, ref = ctxs(name)
    N = be.params.N
    rng = np.random.default_rng(1)
    polys = rng.integers(0, GOLDILOCKS_P, (5, N), dtype=np.uint64)
    polys[0] = 0; polys[0, 1] = 1
    fwd = be.debug_ntt(polys)
    for i in range(len(polys)):
        assert np.array_equal(fwd[i], ref.ntt(polys[i])), f"forward NTT poly {i}"
    inv = be.debug_ntt(fwd, inverse=True)
    assert np.array_equal(inv, polys)


@pytest.mark.parametrize("name", TOYS)
def test_keys_bit_exact(ctxs, name):
    be, ref = ctxs(name)
    g, r = be.debug_keys(), ref.keys()
    for what, a, b in zip(("s_lwe", "s_big", "ksk", "bsk_coef"), g, r):
        assert np.array_equal(a, b), what


@pytest.mark.parametrize("name", TOYS)
def test_encrypt_decrypt_bit_exact(ctxs, name):
    be, ref = ctxs(name)
    p = 7
    msgs = np.arange(-3, 2 * p + 3, dtype=np.int32)
    ids = np.arange(len(msgs), dtype=np.uint64) * 3 + 11
    g = be.debug_encrypt(p, msgs, ids, enc_seed=99)
    r = ref.encrypt(p, msgs, ids, 99)
    assert np.array_equal(g, r)
    assert np.array_equal(be.debug_decrypt(p, g), msgs % (2 * p))
    assert np.array_equal(be.debug_decrypt(p, g), ref.decrypt(p, r))


@pytest.mark.parametrize("name", TOYS)
@pytest.mark.parametrize("p", [2, 3, 5, 7])
def test_pbs_every_stage_bit_exact(ctxs, name, p):
    be, ref = ctxs(name)
    rng = np.random.default_rng(100 + p)
    for tab, mode in tables_for(p, rng):
        L = len(tab)
        msgs = np.arange(L, dtype=np.int32)
        cts = ref.encrypt(p, msgs, np.arange(L) + 5, 42)
        tables = np.zeros((L, 2 * p), np.uint8); tables[:, :L] = tab
        out, ks, ms, acc = be.debug_pbs(p, cts, tables, np.full(L, L, np.uint8), np.full(L, mode, np.int32))
        for m in range(L):
            ro, rks, rms, racc = ref.pbs(p, cts[m], tab, mode)
            assert np.array_equal(ks[m], rks), f"key switch m={m}"
            assert np.array_equal(ms[m], rms), f"modulus switch m={m}"
            assert np.array_equal(acc[m], racc), f"blind rotation m={m}"
            assert np.array_equal(out[m], ro), f"sample extract m={m}"
        assert np.array_equal(be.debug_decrypt(p, out), np.array(tab)), (tab, mode)


@pytest.mark.parametrize("circuit,p,toy", [("half_adder", 15, "toy3"), ("full_adder", 11, "toy3"), ("_2_input_gates", 15, "toy5"),
                                           ("ascon_lut", 11, "toy3"), ("aes_sbox", 11, "toy5"), ("trivium_iter_v1", 15, "toy5")])
def test_program_equals_oracle_and_cleartext(ctxs, circuit, p, toy):
    be, ref = ctxs(toy)
    e = next(x for x in load_ref_mapped() if x["circuit"] == circuit and x["p"] == p and x["mapper"] == "search" and not x.get("strict"))
    env = read_lbf(e["lbf"])
    prog = levelize(env, p)
    cp = be.load(prog)
    B = 40
    inputs = selfcheck_inputs(e["input_names"])
    bits = np.array([inputs[nm][:B] for nm in prog.input_names], dtype=np.uint8)
    got = be.eval_bits(cp, bits)
    seed0 = be.last_enc_seed                     # every encrypting call draws a fresh seed
    want = unpack_outputs(e, batch=B)
    for nm in prog.output_names:
        assert np.array_equal(got[prog.out_index[nm]], want[str(nm)]), nm
    # ragged chunking (chunk of 16 + 16 + 8) must give the same answer
    CT = be.params.ct_words * 8
    got2 = be.eval_bits(cp, bits, max_wire_bytes=16 * (prog.n_slots * CT + 64 * 1024))
    assert np.array_equal(got, got2)
    if B * prog.n_boots <= 400:
        r = ref.eval_prog(prog, bits[:, :6].copy(), enc_seed=seed0, total=B)
        assert np.array_equal(r, got[:, :6])


def test_drop_in_eval_contract(ctxs):
    """LutExecEnv.eval keeps the reference's return contract (fbs_exec_env.py:208-229): dict name -> int64[B], Python
    scalar for Const outputs, outputs that are lincombs / inputs."""
    be, _ = ctxs("toy3")
    from tfhe_fbs_map_b200 import LutExecEnv
    env = LutExecEnv()
    a, b = env.input("a"), env.input("b")
    d = env.linear([1, 2], [a, b]); e = env.linear([1, 1], [env.const(1), d]); f = env.bootstrap(e, [1, 0, 1, 1, 0])
    g = env.linear([2, 1], [a, f]); h = env.bootstrap(g, [1, 1, 0, 2])
    env.output("f", f); env.output("g", g); env.output("h", h); env.output("one", env.const(1)); env.output("a", a)
    iv = {"a": [1, 0, 1, 0], "b": [1, 0, 0, 1]}
    got = env.eval(iv, backend=be)
    want = cleartext.lut_eval(env, iv)
    assert got["one"] == 1 and isinstance(got["one"], int)
    for k in ("f", "g", "h", "a"):
        assert got[k].dtype == np.int64 and np.array_equal(got[k], want[k]), k
    clear = env.eval_clear(iv, backend=be)
    for k in ("f", "g", "h", "a"):
        assert np.array_equal(clear[k], want[k])


@pytest.mark.parametrize("name", sorted({e["circuit"] for e in load_ref_mapped()}))
def test_bit_env_eval_on_gpu_matches_reference_outputs(ctxs, name):
    be, _ = ctxs("toy1")
    entry = next(e for e in load_ref_mapped() if e["circuit"] == name)
    env = read_golden_blif(name)
    got = env.eval(selfcheck_inputs(entry["input_names"]), backend=be)
    want = unpack_outputs(entry)
    for k in got:
        assert np.array_equal(np.asarray(got[k]), want[str(k)]), k


@pytest.mark.parametrize("item", load_lbf_index(), ids=lambda e: e["file"])
def test_clear_kernel_reproduces_reference_hashes(ctxs, item):
    be, _ = ctxs("toy1")
    env = read_golden_lbf(item["file"])
    got = env.eval_clear(selfcheck_inputs(item["input_names"]), backend=be)
    assert out_hash(got) == item["out_sha256"]


def test_error_paths(ctxs):
    from tfhe_fbs_map_b200.backend import B200Backend, FbsError
    be, _ = ctxs("toy1")
    nk = B200Backend("toy1", device=0, keygen=False)
    env = read_golden_lbf("adder8_p15.lbf")
    with pytest.raises(FbsError):
        nk.eval_bits(nk.load(levelize(env, 15)), np.zeros((16, 4), np.uint8))
    with pytest.raises(AssertionError):
        levelize(env, 5)                       # tables of 13 entries do not fit p=5
    nk.close()


def test_malformed_programs_and_unsafe_sub_ranges_are_refused(ctxs):
    """fbs_prog_load validates every CSR / level pointer array and the table values before anything is allocated, and
    fbs_run_level refuses a node sub-range on a program whose slots are recycled (it would corrupt live wires)."""
    import ctypes
    import torch
    from tfhe_fbs_map_b200.backend import FbsError
    be, _ = ctxs("toy3")
    env = read_golden_lbf("adder8_p15.lbf")
    prog = levelize(env, 15)                                   # reuse_slots=True: NOT contiguous
    cp = be.load(prog)
    B = 2
    wires = torch.empty(be.wires_bytes(cp, B) // 8, dtype=torch.int64, device="cuda")
    d_in = torch.zeros((prog.n_inputs, B), dtype=torch.uint8, device="cuda")
    be.encrypt_inputs(cp, d_in.data_ptr(), B, wires.data_ptr())
    width0 = int(prog.arrays["bs_level_ptr"][1] - prog.arrays["bs_level_ptr"][0])
    assert width0 >= 2
    with pytest.raises(FbsError, match="contiguous_levels"):
        be.run_level(cp, 0, B, wires.data_ptr(), 0, 1)
    be.run_level(cp, 0, B, wires.data_ptr(), 0, width0)        # the whole level by explicit range is fine
    be.run_level(cp, 0, B, wires.data_ptr())
    torch.cuda.synchronize()

    def broken(**patch):
        p2 = levelize(env, 15)
        for k, f in patch.items():
            p2.arrays[k] = np.ascontiguousarray(f(p2.arrays[k].copy()))
        return p2

    def bump_last(a):
        a[-1] += 1
        return a

    def descending(a):
        a[1], a[2] = a[2] + 5, a[1]
        return a

    def big_entry(a):
        a[0] = 200
        return a

    def bad_mode(a):
        a[0] = -1
        return a
    for patch in (dict(lc_ptr=descending), dict(bs_tab_ptr=descending), dict(bs_level_ptr=bump_last), dict(lc_level_ptr=bump_last),
                  dict(out_ptr=descending), dict(bs_tab=big_entry), dict(bs_mode=bad_mode)):
        with pytest.raises(FbsError):
            be.load(broken(**patch))


def test_fused_peer_store_epilogue_single_gpu(ctxs):
    """The sample-extract epilogue writes each output ciphertext into every registered peer replica.  With one GPU the
    'peers' are two more buffers on the same device: after a run all three must hold identical bootstrap outputs."""
    import ctypes
    import torch
    be, _ = ctxs("toy3")
    e = next(x for x in load_ref_mapped() if x["circuit"] == "ascon_lut" and x["p"] == 11 and x["mapper"] == "search")
    prog = levelize(read_lbf(e["lbf"]), 11, shard_pad=2)
    cp = be.load(prog)
    B = 8
    nbytes = be.wires_bytes(cp, B)
    main_buf, p1, p2, other = (be.wires_alloc(nbytes) for _ in range(4))
    try:
        inputs = selfcheck_inputs(e["input_names"])
        bits = np.array([inputs[nm][:B] for nm in prog.input_names], dtype=np.uint8)
        d_in = torch.from_numpy(bits).cuda()
        for buf in (main_buf, p1, p2, other):               # replicas start from the same encrypted inputs
            be.encrypt_inputs(cp, d_in.data_ptr(), B, buf, enc_seed=5)
        # rank < 0: peer stores only (no device-side flags: the "peers" here are not running anything)
        be.set_peers(main_buf, nbytes, [p1, p2], -1)
        # the binding belongs to main_buf ONLY: a run on any other buffer (fbs_eval_bits / fbs_pbs_batch use internal ones)
        # must not write into the peers
        snap = torch.empty(nbytes // 8, dtype=torch.int64, device="cuda")
        cudart = ctypes.CDLL("libcudart.so")
        cudart.cudaMemcpy(ctypes.c_void_p(snap.data_ptr()), ctypes.c_void_p(p1), ctypes.c_size_t(nbytes), 3)
        be.run(cp, B, other)
        be.pbs_batch(5, np.arange(4, dtype=np.uint8), np.tile(np.array([0, 1, 1, 0, 1, 0, 0, 0, 0, 0], np.uint8), (4, 1)), np.full(4, 5, np.uint8))
        torch.cuda.synchronize()
        after = torch.empty_like(snap)
        cudart.cudaMemcpy(ctypes.c_void_p(after.data_ptr()), ctypes.c_void_p(p1), ctypes.c_size_t(nbytes), 3)
        assert torch.equal(snap, after), "a level on an unregistered buffer stored into the peers"
        be.run(cp, B, main_buf)
        be.set_peers(None, 0, [], 0)
        torch.cuda.synchronize()
        words = nbytes // 8
        host = []
        for buf in (main_buf, p1, p2):
            t = torch.empty(words, dtype=torch.int64, device="cuda")
            ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(buf), ctypes.c_size_t(nbytes), 3)
            host.append(t.cpu().numpy().reshape(prog.n_slots, B, be.params.ct_words))
        a = prog.arrays
        for q in range(prog.n_boots):
            s = int(a["bs_slot"][q])
            assert np.array_equal(host[0][s], host[1][s]) and np.array_equal(host[0][s], host[2][s]), f"bootstrap {q}"
        d_out = torch.empty((len(prog.output_names), B), dtype=torch.uint8, device="cuda")
        be.decrypt_outputs(cp, B, p2, d_out.data_ptr())
        torch.cuda.synchronize()
        want = unpack_outputs(e, batch=B)
        for nm in prog.output_names:
            assert np.array_equal(d_out.cpu().numpy()[prog.out_index[nm]], want[str(nm)]), nm
    finally:
        be.set_peers(None, 0, [], 0)
        for buf in (main_buf, p1, p2, other):
            be.wires_free(buf)


@pytest.mark.parametrize("circuit,multi", [("aes_sbox", False), ("_2_input_gates", True)])
def test_device_side_level_handoff_two_ranks_one_gpu(circuit, multi):
    """Node-sharded levels with the fused peer-store exchange AND the device-side level hand-off (per-level epoch flags,
    no host synchronisation between levels), with both "ranks" living on this one GPU: two backends (same seeded keys), two
    streams, each rank's peer is the other rank's wire buffer.  All levels of both ranks are enqueued without any host
    sync; the result must be bit-identical to a plain one-rank run on the same encrypted inputs."""
    import ctypes
    import torch
    from tfhe_fbs_map_b200.backend import B200Backend
    from tfhe_fbs_map_b200.dist import level_node_range
    e = next(x for x in load_ref_mapped() if x["circuit"] == circuit and x["p"] == 11 and x["mapper"] == "search" and not x.get("strict"))
    prog = levelize(read_lbf(e["lbf"]), 11, shard_pad=2, multi_value=multi)       # multi-value: the sharding unit is the group (one rotation)
    B, world = 4, 2
    bes = [B200Backend("toy3", device=0, seed=21) for _ in range(world)]
    cps = [be.load(prog) for be in bes]
    nbytes = bes[0].wires_bytes(cps[0], B)
    bufs = [be.wires_alloc(nbytes) for be in bes]
    solo = bes[0].wires_alloc(nbytes)
    streams = [torch.cuda.Stream() for _ in range(world)]
    cudart = ctypes.CDLL("libcudart.so")

    def fetch(ptr):
        t = torch.empty(nbytes // 8, dtype=torch.int64, device="cuda")
        cudart.cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes), 3)
        return t.cpu().numpy().reshape(prog.n_slots, B, bes[0].params.ct_words)
    try:
        inputs = selfcheck_inputs(e["input_names"])
        bits = np.array([inputs[nm][:B] for nm in prog.input_names], dtype=np.uint8)
        d_in = torch.from_numpy(bits).cuda()
        for be, cp, buf in zip(bes, cps, bufs):
            be.encrypt_inputs(cp, d_in.data_ptr(), B, buf, enc_seed=9)
        bes[0].encrypt_inputs(cps[0], d_in.data_ptr(), B, solo, enc_seed=9)
        torch.cuda.synchronize()
        for r in range(world):
            bes[r].set_peers(bufs[r], nbytes, [bufs[1 - r]], r)
        a = prog.arrays
        for rep in range(2):                                 # the epochs keep counting across runs
            for lv in range(prog.n_levels):
                width = int(a["grp_level_ptr"][lv + 1] - a["grp_level_ptr"][lv]) if multi else int(a["bs_level_ptr"][lv + 1] - a["bs_level_ptr"][lv])
                for r in range(world):
                    nb, ne, _ = level_node_range(width, world, r)
                    bes[r].run_level(cps[r], lv, B, bufs[r], nb, ne, stream=streams[r].cuda_stream)
            for r in range(world):
                bes[r].run_level_sync(bufs[r], stream=streams[r].cuda_stream)
        torch.cuda.synchronize()
        assert [be.sync_status() for be in bes] == [0, 0]
        for r in range(world):
            bes[r].set_peers(None, 0, [], 0)
        bes[0].run(cps[0], B, solo)
        torch.cuda.synchronize()
        want, got0, got1 = fetch(solo), fetch(bufs[0]), fetch(bufs[1])
        for q in range(prog.n_boots):
            s = int(a["bs_slot"][q])
            assert np.array_equal(want[s], got0[s]) and np.array_equal(want[s], got1[s]), f"bootstrap {q}"
    finally:
        for r in range(world):
            bes[r].set_peers(None, 0, [], 0)
            bes[r].wires_free(bufs[r])
        bes[0].wires_free(solo)
        for be in bes:
            be.close()


def test_edge_cases(ctxs):
    """Batch of one, a circuit without any bootstrap (outputs are inputs / constants / negations only), the smallest
    message space with a full-length negacyclic table, and a ragged batch that is not a multiple of any tile size."""
    from tfhe_fbs_map_b200 import LutExecEnv
    be, _ = ctxs("toy3")
    # no bootstrap at all
    env = LutExecEnv()
    a, b = env.input("a"), env.input("b")
    env.output("na", env.linear([-1], [a], 1)); env.output("b", b); env.output("zero", env.const(0)); env.output("sum", env.linear([1, 1], [a, b]))
    iv = {"a": [0, 1, 1], "b": [1, 1, 0]}
    got = env.eval(iv, fbs_size=3, backend=be)
    assert got["na"].tolist() == [1, 0, 0] and got["b"].tolist() == [1, 1, 0] and got["zero"] == 0 and got["sum"].tolist() == [1, 2, 1]
    # p = 2 with a table of length 2p = 4 (neg mode) and a batch of one
    env = LutExecEnv()
    a, b = env.input("a"), env.input("b")
    x = env.bootstrap(env.linear([1, 2], [a, b]), [0, 1, 1, 0])
    env.output("x", x)
    for va, vb in ((0, 0), (1, 0), (0, 1), (1, 1)):
        assert env.eval({"a": [va], "b": [vb]}, fbs_size=2, backend=be)["x"].tolist() == [[0, 1, 1, 0][va + 2 * vb]]
    # ragged batch
    e = next(x for x in load_ref_mapped() if x["circuit"] == "full_adder" and x["p"] == 15 and x["mapper"] == "search")
    lut = read_lbf(e["lbf"])
    inputs = {k: v[:37] for k, v in selfcheck_inputs(e["input_names"]).items()}
    got = lut.eval(inputs, fbs_size=15, backend=be)
    want = unpack_outputs(e, batch=37)
    for k in got:
        assert np.array_equal(got[k], want[str(k)])
    # a table that is not realisable for the requested p is rejected before anything runs
    env = LutExecEnv()
    a, b, c = env.input("a"), env.input("b"), env.input("c")
    env.output("y", env.bootstrap(env.linear([1, 2, 4], [a, b, c]), [0, 1, 1, 0, 1, 1, 0, 1]))
    with pytest.raises((ValueError, AssertionError)):
        env.eval({"a": [0], "b": [0], "c": [0]}, fbs_size=5, backend=be)
