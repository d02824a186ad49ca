"""Merkle-tree oracle (test infrastructure only).  Restates R/merkle_tree/mod.rs:397-533
(build), :547-636 (proofs), :172-212 and :262-331 (verification), :728-817 (index helpers).

Generic over callables: leaf_hash(leaf) -> leaf digest, convert(leaf digest) -> inner input,
two_to_one(l, r) -> inner digest.  `evaluate` and `compress` coincide for both hash families on
this path (Poseidon: crh/poseidon/mod.rs:58-64; Pedersen: compress serialises then evaluates).
"""
from __future__ import annotations


def tree_height(n: int) -> int:                 # :730-736
    return 1 if n == 1 else n.bit_length()


def parent(i):                                  # :771-778
    return (i - 1) >> 1


def sibling(i):                                 # :752-762
    return i + 1 if i % 2 == 1 else i - 1


class MerkleTree:
    def __init__(self, leaf_digests, bottom_two_to_one, compress):
        """new_with_leaf_digest :424-523.  non_leaf_nodes in heap order (root at 0)."""
        n = len(leaf_digests)
        assert n > 1 and n & (n - 1) == 0, "`leaves.len() should be power of two and greater than one"
        self.leaf_nodes = list(leaf_digests)
        self.height = tree_height(n)
        nodes = [None] * (n - 1)
        start = n // 2 - 1                                                   # bottom non-leaf level
        for i in range(n // 2):
            nodes[start + i] = bottom_two_to_one(leaf_digests[2 * i], leaf_digests[2 * i + 1])
        while start > 0:
            upper = start
            start = (start - 1) // 2
            for cur in range(start, upper):
                nodes[cur] = compress(nodes[2 * cur + 1], nodes[2 * cur + 2])
        self.non_leaf_nodes = nodes
        self._bottom = bottom_two_to_one
        self._compress = compress

    @classmethod
    def new(cls, leaves, leaf_hash, bottom_two_to_one, compress):
        """new :411-422."""
        t = cls([leaf_hash(l) for l in leaves], bottom_two_to_one, compress)
        t._leaf_hash = leaf_hash
        return t

    def root(self):
        return self.non_leaf_nodes[0]

    def generate_proof(self, index: int):
        """:547-575 -> (leaf_sibling_hash, auth_path top->bottom, leaf_index)."""
        n = len(self.leaf_nodes)
        sib = self.leaf_nodes[index ^ 1]
        cur = parent(index + n - 1)
        path = []
        while cur != 0:
            path.append(self.non_leaf_nodes[sibling(cur)])
            cur = parent(cur)
        path.reverse()
        return (sib, path, index)

    def generate_multi_proof(self, indexes):
        """:589-623: sorted/deduped indexes, front-incremental encoding of the auth paths."""
        idx = sorted(set(indexes))
        prefix, suffixes, sibs = [], [], []
        prev = []
        for i in idx:
            sib, path, _ = self.generate_proof(i)
            sibs.append(sib)
            k = 0
            while k < len(prev) and k < len(path) and prev[k] == path[k]:
                k += 1
            prefix.append(k)
            suffixes.append(path[k:])
            prev = path
        return (sibs, prefix, suffixes, idx)


def verify_path(proof, leaf_digest, root, bottom_two_to_one, compress) -> bool:
    """Path::verify :172-212 (leaf already hashed by the caller)."""
    sib, path, index = proof
    l, r = (leaf_digest, sib) if index & 1 == 0 else (sib, leaf_digest)
    cur = bottom_two_to_one(l, r)
    index >>= 1
    for level in range(len(path) - 1, -1, -1):
        l, r = (cur, path[level]) if index & 1 == 0 else (path[level], cur)
        cur = compress(l, r)
        index >>= 1
    return cur == root


def verify_multi_path(multi_proof, leaf_digests, root, bottom_two_to_one, compress) -> bool:
    """MultiPath::verify :262-331, sequential, with the reference's look-up table of already hashed nodes
    (`hash_lut.entry(index_in_tree).or_insert_with(..)`, :304-306 and :316-318): a node is hashed by the FIRST path that
    reaches it; later paths take the table's value and never compare their own (left, right) with it.
    multi_proof = (leaf_siblings_hashes, prefix_lengths, suffixes, leaf_indexes); leaf_digests already hashed."""
    sibs, prefix, suffixes, idx = multi_proof
    height = len(suffixes[0]) + 2                          # :269
    lut = {}
    prev = list(suffixes[0])                               # :277
    for i, leaf_index in enumerate(idx):
        k = prefix[i]
        auth = list(suffixes[i]) if k == 0 else prev[:k] + list(suffixes[i])    # prefix_decode_path :284-288
        prev = auth
        claimed, sib = leaf_digests[i], sibs[i]
        l, r = (claimed, sib) if leaf_index & 1 == 0 else (sib, claimed)
        index = leaf_index >> 1
        in_tree = parent(leaf_index + (1 << (height - 1)) - 1)
        if in_tree not in lut:
            lut[in_tree] = bottom_two_to_one(l, r)
        cur = lut[in_tree]
        for level in range(len(auth) - 1, -1, -1):
            l, r = (cur, auth[level]) if index & 1 == 0 else (auth[level], cur)
            index >>= 1
            in_tree = parent(in_tree)
            if in_tree not in lut:
                lut[in_tree] = compress(l, r)
            cur = lut[in_tree]
        if cur != root:                                     # :321-323
            return False
    return True
