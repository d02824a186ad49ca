"""Property tests (hypothesis) of the host-side pieces that have no GPU in them: wire-format round trips and the Absorb
encodings of the mirror against the oracle's restatement, on generated nested values."""
import numpy as np
from hypothesis import given, settings, strategies as st

import crypto_primitives_b200 as cp
from crypto_primitives_b200 import serialize as S
from crypto_primitives_b200.merkle_tree import MultiPath, Path
from crypto_primitives_b200.sponge import absorb as A
from oracle import absorb as OA

F = cp.BLS12_381_FR
P = F.modulus
felem = st.integers(min_value=0, max_value=P - 1)


@settings(max_examples=60, deadline=None)
@given(sib=felem, auth=st.lists(felem, max_size=24), idx=st.integers(min_value=0, max_value=2**40))
def test_path_round_trip(sib, auth, idx):
    codec = S.FieldDigest(F)
    p = Path(F.elements([sib])[0], [F.elements([a])[0] for a in auth], idx)
    b = S.ser_path(p, codec)
    assert len(b) == 32 + 8 + 32 * len(auth) + 8
    q = S.de_path(b, codec)
    assert q.leaf_index == idx and F.to_ints(q.leaf_sibling_hash) == [sib]
    assert (F.to_ints(np.array(q.auth_path)) if auth else []) == auth
    assert S.ser_path(q, codec) == b


@settings(max_examples=40, deadline=None)
@given(data=st.data())
def test_multipath_round_trip(data):
    k = data.draw(st.integers(min_value=1, max_value=6))
    sibs = data.draw(st.lists(felem, min_size=k, max_size=k))
    pre = data.draw(st.lists(st.integers(min_value=0, max_value=30), min_size=k, max_size=k))
    suf = [data.draw(st.lists(felem, max_size=5)) for _ in range(k)]
    idx = sorted(data.draw(st.lists(st.integers(min_value=0, max_value=2**30), min_size=k, max_size=k)))
    codec = S.FieldDigest(F)
    mp = MultiPath([F.elements([s])[0] for s in sibs], pre, [[F.elements([x])[0] for x in s] for s in suf], idx)
    b = S.ser_multipath(mp, codec)
    m2 = S.de_multipath(b, codec)
    assert m2.auth_paths_prefix_lenghts == pre and m2.leaf_indexes == idx
    assert [[F.to_ints(x)[0] for x in s] for s in m2.auth_paths_suffixes] == suf
    assert S.ser_multipath(m2, codec) == b


# a generated absorbable, built in parallel for the mirror (A) and the oracle (OA)
def absorbables():
    leaf = st.one_of(
        st.booleans().map(lambda v: (v, v)),
        st.integers(min_value=-2**31, max_value=2**31 - 1).map(lambda v: (v, v)),
        st.tuples(st.integers(min_value=0, max_value=2**64 - 1), st.sampled_from([64, 128])).map(lambda t: (A.UInt(t[0], t[1]), OA.UInt(t[0], t[1]))),
        st.tuples(st.integers(min_value=-2**63, max_value=2**63 - 1)).map(lambda t: (A.SInt(t[0], 64), OA.SInt(t[0], 64))),
        st.binary(max_size=80).map(lambda b: (b, b)),
        st.text(max_size=12).map(lambda s: (s, s)),
        felem.map(lambda v: (A.Elems(F, F.elements([v])), OA.Fe(v, P))),
        st.none().map(lambda _: (None, None)),
    )
    def extend(children):
        return st.one_of(
            st.lists(children, max_size=4).map(lambda xs: ([x[0] for x in xs], [x[1] for x in xs])),
            children.map(lambda x: (A.Some(x[0]), OA.Some(x[1]))),
            st.binary(max_size=20).map(lambda b: (A.WithLength(b), OA.WithLength(b))),
        )
    return st.recursive(leaf, extend, max_leaves=8)


@settings(max_examples=150, deadline=None)
@given(pair=absorbables())
def test_absorb_encodings_match_the_oracle(pair):
    g, o = pair
    assert A.to_sponge_bytes(g) == OA.to_sponge_bytes(o)
    assert F.to_ints(A.to_sponge_field_elements(g, F)) == OA.to_sponge_field_elements(o, P)
