"""Oracle self-checks for the Pedersen / Jubjub path.  PARITY UNPINNED: the reference holds no
golden vectors here (only native==gadget under an RNG that cannot be replayed without Rust), so
these tests tie the two restatements together and check the algebra the reference relies on."""
import numpy as np

from oracle import cref, fields as OF, jubjub as jj, pedersen as PD

W = PD.Window(4, 256)                                   # R/merkle_tree/tests/mod.rs:12-17


def test_curve_constants():
    assert jj.D == 19257038036680949359750312669786877991949435402254120286184196891950884077233
    assert pow(jj.A, (jj.Q - 1) // 2, jj.Q) == 1        # a square
    assert pow(jj.D, (jj.Q - 1) // 2, jj.Q) == jj.Q - 1  # d non-square  => complete addition law
    G = (8076246640662884909881801758704306714034609987455869804520522091855516602923,
         13262374693698910701929044844600465831413122818447359594527400194675274060458)
    assert jj.is_on_curve(G) and jj.mul(jj.ORDER, G) == jj.IDENTITY and jj.mul(jj.COFACTOR, G) != jj.IDENTITY


def test_group_law():
    rng = OF.SplitMix64(3)
    Pt, Qt, Rt = (PD.synthetic_base(rng) for _ in range(3))
    assert jj.add(Pt, jj.IDENTITY) == Pt
    assert jj.add(Pt, jj.neg(Pt)) == jj.IDENTITY
    assert jj.add(jj.add(Pt, Qt), Rt) == jj.add(Pt, jj.add(Qt, Rt))
    assert jj.double(Pt) == jj.mul(2, Pt) and jj.is_on_curve(jj.add(Pt, Qt))
    assert jj.mul(jj.ORDER, Pt) == jj.IDENTITY


def test_crh_is_linear_in_bits_and_pads_with_zeros():
    prm = PD.setup(PD.Window(4, 8), 5)
    w = PD.Window(4, 8)
    a = PD.crh_evaluate(prm, w, bytes([0b0101, 0, 0, 0]))
    b = PD.crh_evaluate(prm, w, bytes([0b1010, 0, 0, 0]))
    assert jj.add(a, b) == PD.crh_evaluate(prm, w, bytes([0b1111, 0, 0, 0]))
    assert PD.crh_evaluate(prm, w, bytes([7])) == PD.crh_evaluate(prm, w, bytes([7, 0, 0, 0]))
    assert PD.crh_evaluate(prm, w, b"") == jj.IDENTITY
    # window value * G_w view (generators are successive doublings, crh/pedersen/mod.rs:48-56)
    assert PD.crh_evaluate(prm, w, bytes([0x0b])) == jj.mul(0x0b, prm.generators[0][0])


def test_c_restatement_equals_python():
    prm = PD.setup(W, 7, commitment=True)
    c = cref.Pedersen(prm, W)
    inp = cref.synth_bytes(9, 3 * 128).reshape(3, 128)
    out = c.batch(inp)
    for i in range(3):
        assert tuple(cref.mont_to_ints(out[i], jj.Q)) == PD.crh_evaluate(prm, W, bytes(inp[i]))
    short = np.ascontiguousarray(inp[:, :32])                        # 32-byte leaves, as in R/merkle_tree/tests/mod.rs:96
    rs = [OF.SplitMix64(40 + i).field(OF.JUBJUB_FR) for i in range(3)]
    r = np.stack([np.frombuffer(x.to_bytes(32, "little"), dtype=np.uint8) for x in rs])
    out = c.batch(short, r)
    for i in range(3):
        assert tuple(cref.mont_to_ints(out[i], jj.Q)) == PD.commit(prm, W, bytes(short[i]), rs[i])
    # two-to-one compress == CRH over x||y bytes of both children (crh/pedersen/mod.rs:187-197)
    kids = c.batch(inp)[:2][None]
    got = c.compress_batch(kids)[0]
    pts = [tuple(cref.mont_to_ints(kids[0, k], jj.Q)) for k in range(2)]
    assert tuple(cref.mont_to_ints(got, jj.Q)) == PD.two_to_one_compress(prm, W, pts[0], pts[1])


def test_c_pedersen_merkle_small():
    prm = PD.setup(W, 11)
    c = cref.Pedersen(prm, W)
    leaves = cref.synth_bytes(12, 4 * 32).reshape(4, 32)
    ln, nn = cref.pedersen_merkle(c, c, leaves, threads=2)
    d = [PD.crh_evaluate(prm, W, bytes(l)) for l in leaves]
    n1 = PD.two_to_one_compress(prm, W, d[0], d[1])
    n2 = PD.two_to_one_compress(prm, W, d[2], d[3])
    root = PD.two_to_one_compress(prm, W, n1, n2)
    assert [tuple(cref.mont_to_ints(x, jj.Q)) for x in nn] == [root, n1, n2]


def test_bowe_hopwood_c_equals_python_and_encoding():
    """R/crh/bowe_hopwood/mod.rs:115-185: chunk value v -> (1 + c0 + 2*c1) * (-1)^c2 times its generator; only covered
    chunks contribute; the two restatements agree."""
    from oracle import bowe_hopwood as OBH
    w = PD.Window(63, 2)
    prm = OBH.setup(w, 3)
    g = prm.generators[0]
    for v, e in ((0, 1), (1, 2), (2, 3), (3, 4), (4, -1), (5, -2), (6, -3), (7, -4)):
        pt = jj.add(jj.add(jj.mul(e % jj.ORDER, g[0]), g[1]), g[2])          # a 1-byte input covers chunks 0, 1, 2
        assert OBH.crh_evaluate(prm, w, bytes([v])) == pt[0]
    assert OBH.crh_evaluate(prm, w, b"") == 0
    c = cref.Pedersen(prm, w)
    for ln in (0, 1, 20, 47):
        inp = np.ascontiguousarray(cref.synth_bytes(5 + ln, 3 * max(ln, 1)).reshape(3, max(ln, 1))[:, :ln])
        out = c.bowe_hopwood_batch(inp)
        for i in range(3):
            assert cref.mont_to_ints(out[i], jj.Q)[0] == OBH.crh_evaluate(prm, w, bytes(inp[i]))
