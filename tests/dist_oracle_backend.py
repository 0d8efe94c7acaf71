"""A CPU backend for stark-rs_b200/distributed.py built on the ORACLE (tests only): lets the sharding, gather and
proof-assembly logic of the sharded prover run on CPU with the gloo backend.  TEST INFRASTRUCTURE."""
import numpy as np
import torch

import oracle as O

P = 998244353


class _Tree:
    def __init__(self, nodes, n):
        self.nodes, self.n = nodes, n              # (2n-1, 32) uint8, leaves first (merkle.rs:18-29)
        self.depth = n.bit_length() - 1

    @property
    def root(self):
        return torch.from_numpy(self.nodes[-1].copy())

    def root_bytes(self):
        return self.nodes[-1].tobytes()

    def open_batch(self, idx):
        out = np.zeros((len(idx), self.depth, 32), dtype=np.uint8)
        for k, i in enumerate(idx):
            out[k] = O.merkle_open(self.nodes[: self.n], int(i))
        return out

    def free(self):
        pass


class OracleBackend:
    device = "cpu"

    def new_codeword(self, n):
        return torch.zeros(n, dtype=torch.int32)

    def new_hashes(self, n):
        return torch.zeros((n, 32), dtype=torch.uint8)

    def upload(self, values):
        return torch.from_numpy(np.asarray(values, dtype=np.uint64).astype(np.int64).astype(np.int32))

    def download(self, cw):
        return (cw.numpy().astype(np.int64) & 0xFFFFFFFF).astype(np.uint64)

    def subtree(self, cw, lo, cnt, width=1):
        v = self.download(cw[lo:lo + cnt * width])
        if width > 1:                                   # column-major [width][cnt] -> row-major leaves
            v = np.ascontiguousarray(v.reshape(width, cnt).T).reshape(-1)
        return _Tree(O.merkle_build(O.hash_leaves(v, width)), cnt)

    def tree_from_hashes(self, hashes):
        return _Tree(O.merkle_build(hashes.numpy()), hashes.shape[0])

    def fold_range(self, cw, n, alpha_raw, offset, omega, lo, cnt, out):
        full = O.fri_fold(self.download(cw[:n]), alpha_raw, offset, omega)   # the oracle folds everything; keep our range
        out[lo:lo + cnt] = self.upload(full[lo:lo + cnt])

    def gather_values(self, cw, idx):
        v = self.download(cw)
        return [int(v[i]) for i in idx]

    def lde(self, cols, n_cols, log_n, log_blowup, offset):
        v = self.download(cols).reshape(n_cols, 1 << log_n)
        return self.upload(np.concatenate([O.fast_lde(c, log_n, log_blowup, offset) for c in v]))

    def challenge(self, transcript):
        return O.fs_challenge(bytes(transcript))

    def hash_from_u64(self, v):
        return O.hash_from_u64(v)

    def sample_indices(self, seed, size, reduced, number):
        return [int(x) for x in O.fri_sample_indices(seed, size, reduced, number)]

    def num_rounds(self, n, ef, nq):
        return O.fri_num_rounds(n, ef, nq)

    def prim_nth_root(self, n):
        return O.ff_prim_nth_root(n)
