"""merkle_tree::{Config, MerkleTree, Path, MultiPath} -- host mirror of R/merkle_tree/mod.rs over
the CUDA library (R = /root/reference/crypto-primitives/src).

The build (MerkleTree::new, mod.rs:411-523) runs on the GPU: one kernel hashes all leaves, then one
launch per level compresses contiguous child pairs of the heap-ordered node array.  The tree object
keeps the reference's two arrays -- leaf_nodes[n] and non_leaf_nodes[n-1] (root at 0, children of i
at 2i+1 / 2i+2, mod.rs:383-395) -- on the host, so proofs are index arithmetic exactly as in
mod.rs:547-623; verification and update re-hash through the same GPU entry points.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _native as N
from .crh import pedersen as pedersen_crh
from .crh import poseidon as poseidon_crh


def _p(a):
    return a.ctypes.data_as(N.u64p)


class Config:
    """merkle_tree::Config (mod.rs:83-122) as a value: which leaf hash, converter and two-to-one hash.
    Subclasses provide batched GPU evaluation of the three associated functions."""

    digest_words = 4            # uint64 words per inner digest

    def leaf_hash_batch(self, leaf_param, leaves, device):
        raise NotImplementedError

    def two_to_one_batch(self, param, pairs, device):
        """pairs: (n, 2, digest_words) of already converted digests -> (n, digest_words)."""
        raise NotImplementedError

    def build(self, leaf_param, two_to_one_param, leaves, device):
        """Generic build (mod.rs:411-523): all leaves in one GPU batch, then one GPU batch per level.  Configs with a
        fused device-side build (`cpb_merkle_*_build`) override this; the generic form round-trips each level
        through the host."""
        leaf_nodes = self.leaf_hash_batch(leaf_param, leaves, device)
        return leaf_nodes, self.build_from_digests(two_to_one_param, leaf_nodes, device)

    def _pairs(self, nodes):
        return np.ascontiguousarray(nodes).reshape((-1, 2) + tuple(nodes.shape[1:]))

    def build_from_digests(self, two_to_one_param, leaf_digests, device):
        d = np.ascontiguousarray(leaf_digests, dtype=np.uint64)
        n = d.shape[0]
        if n < 2 or n & (n - 1):
            raise ValueError("`leaves.len() should be power of two and greater than one")
        nodes = np.empty((n - 1,) + d.shape[1:], dtype=np.uint64)
        start = n // 2 - 1
        nodes[start:] = self.two_to_one_batch(two_to_one_param, self._pairs(d), device)
        while start > 0:
            upper, start = start, (start - 1) // 2
            nodes[start:upper] = self.two_to_one_batch(two_to_one_param, self._pairs(nodes[upper:2 * upper + 1]), device)
        return nodes

    def default_leaf_digest(self):
        return np.zeros(self.digest_words, dtype=np.uint64)


class PoseidonFieldConfig(Config):
    """Config{Leaf=[F], LeafDigest=InnerDigest=F, IdentityDigestConverter, LeafHash=poseidon::CRH,
    TwoToOneHash=poseidon::TwoToOneCRH} -- FieldMTConfig of R/merkle_tree/tests/mod.rs:198-206."""

    def leaf_hash_batch(self, leaf_param, leaves, device):
        return poseidon_crh.CRH.evaluate_batch(leaf_param, leaves, device)

    def two_to_one_batch(self, param, pairs, device):
        return poseidon_crh.TwoToOneCRH.compress_batch(param, pairs, device)

    def build(self, leaf_param, two_to_one_param, leaves, device):
        lv = np.ascontiguousarray(leaves, dtype=np.uint64)
        assert lv.ndim == 3 and lv.shape[2] == 4, "leaves must be (n, leaf_len, 4)"
        n, ln = lv.shape[0], lv.shape[1]
        leaf_nodes = np.empty((n, 4), dtype=np.uint64)
        non_leaf = np.empty((max(n - 1, 0), 4), dtype=np.uint64)
        N.check(N.lib.cpb_merkle_poseidon_build(leaf_param.context(device), two_to_one_param.context(device),
                                                _p(lv), ln, n, _p(leaf_nodes), _p(non_leaf)))
        return leaf_nodes, non_leaf

    def build_from_digests(self, two_to_one_param, leaf_digests, device):
        d = np.ascontiguousarray(leaf_digests, dtype=np.uint64).reshape(-1, 4)
        n = d.shape[0]
        non_leaf = np.empty((max(n - 1, 0), 4), dtype=np.uint64)
        N.check(N.lib.cpb_merkle_poseidon_from_digests(two_to_one_param.context(device), _p(d), n, _p(non_leaf)))
        return non_leaf


class PedersenByteConfig(Config):
    """Config{Leaf=[u8], LeafDigest=InnerDigest=C::Affine, ByteDigestConverter, LeafHash=pedersen::CRH,
    TwoToOneHash=pedersen::TwoToOneCRH} -- JubJubMerkleTreeParams of R/merkle_tree/tests/mod.rs:19-33.
    Digests are affine points, (2, 4) words."""

    digest_words = 8

    def leaf_hash_batch(self, leaf_param, leaves, device):
        return pedersen_crh.CRH.evaluate_batch(leaf_param, leaves, device)

    def two_to_one_batch(self, param, pairs, device):
        return pedersen_crh.TwoToOneCRH.compress_batch(param, np.asarray(pairs).reshape(-1, 2, 2, 4), device)

    def build(self, leaf_param, two_to_one_param, leaves, device):
        lv = np.ascontiguousarray(leaves, dtype=np.uint8)
        assert lv.ndim == 2, "leaves must be (n, leaf_len) bytes"
        n, ln = lv.shape
        leaf_nodes = np.empty((n, 2, 4), dtype=np.uint64)
        non_leaf = np.empty((max(n - 1, 0), 2, 4), dtype=np.uint64)
        try:
            N.check(N.lib.cpb_merkle_pedersen_build(leaf_param.context(device), two_to_one_param.context(device),
                                                    lv.ctypes.data_as(N.u8p), ln, n, _p(leaf_nodes), _p(non_leaf)))
        except N.CpbError as e:
            if e.status == N.CPB_BAD_LENGTH:
                raise ValueError("incorrect input length") from e
            raise
        return leaf_nodes, non_leaf

    def build_from_digests(self, two_to_one_param, leaf_digests, device):
        return Config.build_from_digests(self, two_to_one_param, np.asarray(leaf_digests, dtype=np.uint64).reshape(-1, 2, 4), device)

    def default_leaf_digest(self):
        raise NotImplementedError("Affine::default() (the identity) is not representable as an input digest here")


class PedersenPoseidonConfig(Config):
    """Config{Leaf=[u8], LeafHash=PedersenCRHCompressor<C,TECompressor,W>, LeafDigest=InnerDigest=Fq,
    IdentityDigestConverter, TwoToOneHash=poseidon::TwoToOneCRH<Fq>} (BASELINE config 5)."""

    def leaf_hash_batch(self, leaf_param, leaves, device):
        return pedersen_crh.PedersenCRHCompressor.evaluate_batch(leaf_param, leaves, device)

    def two_to_one_batch(self, param, pairs, device):
        return poseidon_crh.TwoToOneCRH.compress_batch(param, pairs, device)

    def build(self, leaf_param, two_to_one_param, leaves, device):
        lv = np.ascontiguousarray(leaves, dtype=np.uint8)
        n, ln = lv.shape
        leaf_nodes = np.empty((n, 4), dtype=np.uint64)
        non_leaf = np.empty((max(n - 1, 0), 4), dtype=np.uint64)
        try:
            N.check(N.lib.cpb_merkle_mixed_build(leaf_param.context(device), two_to_one_param.context(device),
                                                 lv.ctypes.data_as(N.u8p), ln, n, _p(leaf_nodes), _p(non_leaf)))
        except N.CpbError as e:
            if e.status == N.CPB_BAD_LENGTH:
                raise ValueError("incorrect input length") from e
            raise
        return leaf_nodes, non_leaf

    def build_from_digests(self, two_to_one_param, leaf_digests, device):
        return PoseidonFieldConfig().build_from_digests(two_to_one_param, leaf_digests, device)


class BoweHopwoodByteConfig(Config):
    """Config{Leaf=[u8], LeafHash=bowe_hopwood::CRH, LeafDigest=InnerDigest=Fq (the x-coordinate),
    IdentityDigestConverter, TwoToOneHash=bowe_hopwood::TwoToOneCRH} -- the Zcash-style tree (R/crh/bowe_hopwood/mod.rs:
    112-241).  Hashing runs on the GPU (one batch per level); the level loop is the generic host one."""

    def leaf_hash_batch(self, leaf_param, leaves, device):
        from .crh import bowe_hopwood as bh
        return bh.CRH.evaluate_batch(leaf_param, np.ascontiguousarray(leaves, dtype=np.uint8), device)

    def two_to_one_batch(self, param, pairs, device):
        from .crh import bowe_hopwood as bh
        return bh.TwoToOneCRH.compress_batch(param, np.asarray(pairs, dtype=np.uint64).reshape(-1, 2, 4), device)


# ---- index helpers, mod.rs:728-786
def tree_height(num_leaves: int) -> int:
    return 1 if num_leaves == 1 else num_leaves.bit_length()


def is_root(i):
    return i == 0


def left_child(i):
    return 2 * i + 1


def right_child(i):
    return 2 * i + 2


def sibling(i):
    if i == 0:
        return None
    return i + 1 if is_left_child(i) else i - 1


def is_left_child(i):
    return i % 2 == 1


def parent(i):
    return (i - 1) >> 1 if i > 0 else None


def convert_index_to_last_level(index, height):
    return index + (1 << (height - 1)) - 1


@dataclass
class Path:
    """mod.rs:139-152."""
    leaf_sibling_hash: np.ndarray
    auth_path: list
    leaf_index: int

    def verify(self, leaf_hash_params, two_to_one_params, root_hash, leaf, config: Config = None, device: int = 0) -> bool:
        """Path::verify, mod.rs:172-212."""
        cfg = config or PoseidonFieldConfig()
        claimed = cfg.leaf_hash_batch(leaf_hash_params, np.asarray(leaf)[None], device)[0]
        l, r = (claimed, self.leaf_sibling_hash) if self.leaf_index & 1 == 0 else (self.leaf_sibling_hash, claimed)
        cur = cfg.two_to_one_batch(two_to_one_params, np.stack([l, r])[None], device)[0]
        index = self.leaf_index >> 1
        for level in range(len(self.auth_path) - 1, -1, -1):
            sib = self.auth_path[level]
            l, r = (cur, sib) if index & 1 == 0 else (sib, cur)
            cur = cfg.two_to_one_batch(two_to_one_params, np.stack([l, r])[None], device)[0]
            index >>= 1
        return bool(np.array_equal(cur, np.asarray(root_hash, dtype=np.uint64)))


@dataclass
class MultiPath:
    """mod.rs:239-254 (field names as in the reference, including its spelling)."""
    leaf_siblings_hashes: list
    auth_paths_prefix_lenghts: list
    auth_paths_suffixes: list
    leaf_indexes: list

    def verify(self, leaf_hash_params, two_to_one_params, root_hash, leaves, config: Config = None, device: int = 0) -> bool:
        """MultiPath::verify, mod.rs:262-331 (with the same lookup table of already hashed nodes)."""
        cfg = config or PoseidonFieldConfig()
        height = len(self.auth_paths_suffixes[0]) + 2
        lut = {}
        prev = list(self.auth_paths_suffixes[0])
        root = np.asarray(root_hash, dtype=np.uint64)
        claimed_all = cfg.leaf_hash_batch(leaf_hash_params, np.asarray(leaves), device)
        for i, leaf_index in enumerate(self.leaf_indexes):
            k = self.auth_paths_prefix_lenghts[i]
            auth = list(self.auth_paths_suffixes[i]) if k == 0 else prev[:k] + list(self.auth_paths_suffixes[i])
            prev = auth
            claimed, sib = claimed_all[i], self.leaf_siblings_hashes[i]
            l, r = (claimed, sib) if leaf_index & 1 == 0 else (sib, claimed)
            index = leaf_index >> 1
            in_tree = parent(convert_index_to_last_level(leaf_index, height))
            if in_tree not in lut:
                lut[in_tree] = cfg.two_to_one_batch(two_to_one_params, np.stack([l, r])[None], device)[0]
            cur = lut[in_tree]
            for level in range(len(auth) - 1, -1, -1):
                l, r = (cur, auth[level]) if index & 1 == 0 else (auth[level], cur)
                index >>= 1
                in_tree = parent(in_tree)
                if in_tree not in lut:
                    lut[in_tree] = cfg.two_to_one_batch(two_to_one_params, np.stack([l, r])[None], device)[0]
                cur = lut[in_tree]
            if not np.array_equal(cur, root):
                return False
        return True


class MerkleTree:
    """mod.rs:381-395."""

    def __init__(self, config, leaf_nodes, non_leaf_nodes, leaf_hash_param, two_to_one_hash_param, device):
        self.config = config
        self.leaf_nodes = leaf_nodes
        self.non_leaf_nodes = non_leaf_nodes
        self.leaf_hash_param = leaf_hash_param
        self.two_to_one_hash_param = two_to_one_hash_param
        self._height = tree_height(leaf_nodes.shape[0])
        self.device = device

    @classmethod
    def blank(cls, leaf_hash_param, two_to_one_hash_param, height: int, config: Config = None, device: int = 0):
        """mod.rs:400-408: all leaf digests = LeafDigest::default()."""
        cfg = config or PoseidonFieldConfig()
        d = np.tile(cfg.default_leaf_digest(), (1 << (height - 1), 1))
        return cls.new_with_leaf_digest(leaf_hash_param, two_to_one_hash_param, d, cfg, device)

    @classmethod
    def new(cls, leaf_hash_param, two_to_one_hash_param, leaves, config: Config = None, device: int = 0):
        """mod.rs:411-422.  Raises ValueError where the reference asserts (:430-433)."""
        cfg = config or PoseidonFieldConfig()
        try:
            leaf_nodes, non_leaf = cfg.build(leaf_hash_param, two_to_one_hash_param, leaves, device)
        except N.CpbError as e:
            if e.status == N.CPB_NOT_POW2:
                raise ValueError("`leaves.len() should be power of two and greater than one") from e
            raise
        return cls(cfg, leaf_nodes, non_leaf, leaf_hash_param, two_to_one_hash_param, device)

    @classmethod
    def new_with_leaf_digest(cls, leaf_hash_param, two_to_one_hash_param, leaf_digests, config: Config = None, device: int = 0):
        """mod.rs:424-523."""
        cfg = config or PoseidonFieldConfig()
        d = np.ascontiguousarray(leaf_digests, dtype=np.uint64)
        try:
            non_leaf = cfg.build_from_digests(two_to_one_hash_param, d, device)
        except N.CpbError as e:
            if e.status == N.CPB_NOT_POW2:
                raise ValueError("`leaves.len() should be power of two and greater than one") from e
            raise
        return cls(cfg, d, non_leaf, leaf_hash_param, two_to_one_hash_param, device)

    def root(self):
        return self.non_leaf_nodes[0].copy()

    def height(self) -> int:
        return self._height

    def get_leaf_sibling_hash(self, index: int):
        return self.leaf_nodes[index ^ 1].copy()

    def _compute_auth_path(self, index: int):
        """mod.rs:548-573."""
        path = []
        cur = parent(convert_index_to_last_level(index, self._height))
        while not is_root(cur):
            path.append(self.non_leaf_nodes[sibling(cur)].copy())
            cur = parent(cur)
        path.reverse()
        return path

    def generate_proof(self, index: int) -> Path:
        return Path(self.get_leaf_sibling_hash(index), self._compute_auth_path(index), index)

    def generate_multi_proof(self, indexes) -> MultiPath:
        """mod.rs:589-623."""
        idx = sorted(set(int(i) for i in indexes))
        prefix, suffixes, sibs, prev = [], [], [], []
        for i in idx:
            sibs.append(self.get_leaf_sibling_hash(i))
            path = self._compute_auth_path(i)
            k = 0
            while k < len(prev) and k < len(path) and np.array_equal(prev[k], path[k]):
                k += 1
            prefix.append(k)
            suffixes.append(path[k:])
            prev = path
        return MultiPath(sibs, prefix, suffixes, idx)

    def verify_proofs_batch(self, proofs, leaves, root_hash=None) -> np.ndarray:
        """Many Path::verify (mod.rs:172-212) in one kernel launch (field-leaf Config): proofs = list of Path for
        `leaves[i]`; returns a bool array.  One GPU thread recomputes one root."""
        if not isinstance(self.config, PoseidonFieldConfig):
            raise NotImplementedError("batched verification is implemented for the Poseidon field-leaf Config")
        n = len(proofs)
        lv = np.ascontiguousarray(leaves, dtype=np.uint64)
        assert lv.shape[0] == n and lv.ndim == 3
        plen = self._height - 2
        sib = np.ascontiguousarray(np.stack([p.leaf_sibling_hash for p in proofs]), dtype=np.uint64)
        paths = np.ascontiguousarray(np.stack([np.stack(p.auth_path) if plen else np.zeros((0, 4), dtype=np.uint64) for p in proofs]),
                                     dtype=np.uint64).reshape(n, plen, 4)
        idx = np.array([p.leaf_index for p in proofs], dtype=np.uint64)
        root = np.ascontiguousarray(self.root() if root_hash is None else root_hash, dtype=np.uint64)
        ok = np.zeros(n, dtype=np.uint8)
        N.check(N.lib.cpb_merkle_poseidon_verify_batch(self.leaf_hash_param.context(self.device), self.two_to_one_hash_param.context(self.device),
                                                       _p(root), _p(lv), lv.shape[1], _p(sib), _p(paths), plen, _p(idx),
                                                       ok.ctypes.data_as(N.u8p), n))
        return ok.astype(bool)

    def _updated_path(self, index: int, new_leaf):
        """mod.rs:627-677."""
        cfg, dev = self.config, self.device
        new_hash = cfg.leaf_hash_batch(self.leaf_hash_param, np.asarray(new_leaf)[None], dev)[0]
        l, r = (new_hash, self.leaf_nodes[index + 1]) if index & 1 == 0 else (self.leaf_nodes[index - 1], new_hash)
        path = [cfg.two_to_one_batch(self.two_to_one_hash_param, np.stack([l, r])[None], dev)[0]]
        prev = parent(convert_index_to_last_level(index, self._height))
        while not is_root(prev):
            sib = self.non_leaf_nodes[sibling(prev)]
            l, r = (path[-1], sib) if is_left_child(prev) else (sib, path[-1])
            path.append(cfg.two_to_one_batch(self.two_to_one_hash_param, np.stack([l, r])[None], dev)[0])
            prev = parent(prev)
        path.reverse()
        return new_hash, path

    def update(self, index: int, new_leaf):
        """mod.rs:690-701."""
        assert index < self.leaf_nodes.shape[0], "index out of range"
        new_hash, path = self._updated_path(index, new_leaf)
        self.leaf_nodes[index] = new_hash
        cur = convert_index_to_last_level(index, self._height)
        for _ in range(self._height - 1):
            cur = parent(cur)
            self.non_leaf_nodes[cur] = path.pop()

    def check_update(self, index: int, new_leaf, asserted_new_root) -> bool:
        """mod.rs:706-725."""
        assert index < self.leaf_nodes.shape[0], "index out of range"
        new_hash, path = self._updated_path(index, new_leaf)
        if not np.array_equal(path[0], np.asarray(asserted_new_root, dtype=np.uint64)):
            return False
        self.leaf_nodes[index] = new_hash
        cur = convert_index_to_last_level(index, self._height)
        for _ in range(self._height - 1):
            cur = parent(cur)
            self.non_leaf_nodes[cur] = path.pop()
        return True
