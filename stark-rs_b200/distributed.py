"""distributed.py -- the sharded prover: one process per GPU, torch.distributed (NCCL over NVLink) for the
plumbing, the C ABI of libstark_b200.so for every byte of arithmetic.

What shards, and the only two exchanges (north star: "NCCL ... used only to gather subtree roots and folded codewords"):
  * LDE               -- trace columns are independent: rank g owns column groups {g, g+G, ...}; no communication.
  * Merkle commitment -- rank g hashes the contiguous leaf range [g n/G, (g+1) n/G) into a SUBTREE, which is node for node
                         the same as that part of MerkleTree::new's tree (merkle.rs:11-38); all-gather of the G subtree
                         roots (32 G bytes); the top log2 G levels are rebuilt on every rank.
  * FRI fold          -- every rank holds a replica of the current codeword; rank g computes outputs [g h/G, (g+1) h/G) of
                         fold_codeword (fri.rs:57-91) and an all-gather makes the folded codeword replicated again.
  Rounds shorter than `shard_min` are latency-bound and run replicated on every rank (no collective).
The Fiat-Shamir transcript (fiat_shamir.rs), index sampling (fri.rs:176-213) and ProofStream::serialize
(stream.rs:35-64) are host orchestration here; results are byte-identical to the single-GPU path and to the oracle for
every world size (tests/test_distributed_gloo.py on CPU with gloo, tests/test_gpu_distributed.py on the device).

The arithmetic sits behind a small BACKEND interface so the sharding / gather / assembly logic can be exercised on CPU
with the gloo backend (the tests plug the oracle in); the product backend is CudaBackend and has no CPU fallback.
"""
import numpy as np
import torch
import torch.distributed as dist

from .proof_stream import FiatShamir, ProofStream

P = 998244353


# ------------------------------------------------------------------------------------------------- product backend

class _DevView:
    """zero-copy torch view of library-owned device memory"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class _CudaTree:
    def __init__(self, backend, tree):
        self.b, self.t = backend, tree
        self.n = tree.num_leaves
        self.depth = tree.num_levels - 1

    @property
    def root(self):
        """uint8[32] CUDA tensor aliasing the root inside the tree"""
        return torch.as_tensor(_DevView(self.t.root_ptr, 32), device=self.b.device)

    def root_bytes(self):
        return self.t.get_root()

    def open_batch(self, idx):
        return self.t.open_batch(idx) if len(idx) else np.zeros((0, self.depth, 32), dtype=np.uint8)

    def free(self):
        self.t.free()


class CudaBackend:
    """The B200 path: tensors are torch.int32 CUDA tensors holding canonical u32 field elements; all work is queued on
    the context's stream (create the Context on a torch stream and run under `torch.cuda.stream(...)`)."""

    def __init__(self, ctx, device, aux=None):
        """aux: optional second CudaBackend whose Context lives on ANOTHER torch stream (`aux.stream`); lde_commit_sharded
        then builds a group's Merkle tree there while the next group's LDE runs on this one (the NTT is bound by the
        FMA-heavy pipe, the hashing by the ALU pipe: together they fill both)."""
        from . import api
        self.api, self.ctx, self.device = api, ctx, torch.device(device)
        self.aux, self.stream = aux, None

    def _buf(self, t, off=0, n=None):
        n = t.numel() - off if n is None else n
        return self.ctx.wrap(t.data_ptr() + 4 * off, n)

    def new_codeword(self, n):
        return torch.empty(n, dtype=torch.int32, device=self.device)

    def new_hashes(self, n):
        return torch.empty((n, 32), dtype=torch.uint8, device=self.device)

    def upload(self, values):
        v = np.ascontiguousarray(np.asarray(values, dtype=np.uint64))
        if (v >= P).any():
            raise self.api.StarkPanic(1, "non-canonical field element (value >= p) in input")
        return torch.from_numpy(v.astype(np.int64).astype(np.int32)).to(self.device)

    def subtree(self, cw, lo, cnt, width=1):
        """leaf hashes (fri.rs:118-121) + MerkleTree::new over cw[lo:lo+cnt]"""
        b = self._buf(cw, lo, cnt * width)
        t = self.ctx.merkle_build_from_buf(b, cnt, width)
        b.free()
        return _CudaTree(self, t)

    def tree_from_hashes(self, hashes):
        return _CudaTree(self, self.ctx.merkle_build_dev(hashes.data_ptr(), hashes.shape[0]))

    def fold_range(self, cw, n, alpha_raw, offset, omega, lo, cnt, out):
        a, o = self._buf(cw, 0, n), self._buf(out)
        self.ctx.fri_fold_range_dev(a, n, alpha_raw, offset, omega, lo, cnt, o, lo)
        a.free(), o.free()

    def fold_bcast(self, cw, n, alpha_raw, offset, omega, lo, cnt, peer_ptrs, multicast_ptr):
        a = self._buf(cw, 0, n)
        self.ctx.fri_fold_bcast_dev(a, n, alpha_raw, offset, omega, lo, cnt, peer_ptrs, multicast_ptr)
        a.free()

    def gather_values(self, cw, idx):
        if not len(idx):
            return []
        sel = cw[torch.as_tensor(list(idx), dtype=torch.int64, device=self.device)]
        return [int(x) & 0xFFFFFFFF for x in sel.cpu().tolist()]

    def download(self, cw):
        return (cw.cpu().numpy().astype(np.int64) & 0xFFFFFFFF).astype(np.uint64)

    def lde(self, cols, n_cols, log_n, log_blowup, offset):
        out = torch.empty(n_cols << (log_n + log_blowup), dtype=torch.int32, device=self.device)
        a, o = self._buf(cols), self._buf(out)
        self.ctx.lde_dev(a, n_cols, log_n, log_blowup, offset, o)
        a.free(), o.free()
        return out

    def challenge(self, transcript):
        return self.api.fiat_shamir_challenge(transcript)

    def hash_from_u64(self, v):
        return self.api.hash_from_u64(v)

    def sample_indices(self, seed, size, reduced, number):
        return [int(x) for x in self.api.fri_sample_indices(seed, size, reduced, number)]

    def num_rounds(self, n, ef, nq):
        return self.api.fri_num_rounds(n, ef, nq)

    def prim_nth_root(self, n):
        return self.api.prim_nth_root(n)


# ------------------------------------------------------------------------------------------------- collectives

class Comm:
    """the two exchanges of the sharded prover + the path all-reduce of the query phase"""

    def __init__(self, group=None):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.nccl = self.on and dist.get_backend(group) == "nccl"
        self.bytes_gathered = 0

    def all_gather_rows(self, out, mine):
        """out[g] = rank g's `mine`; out is (G, ...) contiguous"""
        if self.world == 1:
            out[0].copy_(mine)
            return
        self.bytes_gathered += out.numel() * out.element_size()
        if self.nccl:
            dist.all_gather_into_tensor(out, mine.contiguous(), group=self.group)
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.contiguous(), group=self.group)
            for g, p in enumerate(parts):
                out[g].copy_(p)

    def all_gather_inplace(self, full, lo, cnt):
        """every rank has written full[lo:lo+cnt] (its own slice, rank-major); make `full` replicated"""
        if self.world == 1:
            return
        self.all_gather_rows(full.view(self.world, cnt), full[lo:lo + cnt].clone())

    def sum_uint8(self, arr, device):
        """element-wise sum over ranks of a host uint8 array (exactly one rank holds each non-zero byte)"""
        if self.world == 1:
            return arr
        t = torch.from_numpy(arr).to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()


class SymmetricArena:
    """One symmetric-memory allocation (torch.distributed._symmetric_memory: every rank's copy is mapped into every other
    rank's address space over NVLink, plus an NVSwitch multicast address when the fabric supports it) from which the
    folded codewords of all FRI rounds are carved, so that the fold kernel can store its results straight into every
    replica (stark_fri_fold_bcast_dev) instead of running a separate all-gather.  NCCL process groups only."""

    def __init__(self, n_elems, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.buf = symm.empty(n_elems, dtype=torch.int32, device=device)
        self.h = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.h.buffer_ptrs]
        mc = 0
        try:
            mc = int(self.h.multicast_ptr or 0)
        except Exception:
            mc = 0
        self.mc = mc
        self.off = 0

    def reset(self):
        self.off = 0

    def carve(self, n):
        """-> (local tensor view, peer addresses, multicast address) of the next n elements"""
        off = self.off
        self.off += (n + 3) & ~3
        assert self.off <= self.buf.numel()
        return (self.buf[off:off + n], [p + 4 * off for p in self.ptrs], (self.mc + 4 * off) if self.mc else 0)

    def barrier(self):
        self.h.barrier(channel=0)


# ------------------------------------------------------------------------------------------------- Merkle

class ShardedTree:
    """MerkleTree (merkle.rs:4-8) of n leaves split into G leaf ranges: `sub` is this rank's subtree (n/G leaves), `top`
    the replicated tree over the G gathered subtree roots.  G = 1: `sub` is the whole tree and `top` is None."""

    def __init__(self, n, sub, top, comm):
        self.n, self.sub, self.top, self.comm = n, sub, top, comm
        self.depth = n.bit_length() - 1
        self.per = n // comm.world if top is not None else n
        self.depth_sub = self.per.bit_length() - 1

    def root_bytes(self):
        return (self.top or self.sub).root_bytes()

    def local_paths(self, idx):
        """(len(idx), depth_sub, 32) uint8: the lower part of MerkleTree::open (merkle.rs:67-80) for the leaves this rank
        owns, zeros for the others"""
        out = np.zeros((len(idx), self.depth_sub, 32), dtype=np.uint8)
        if self.top is None:
            # replicated tree: every rank has it; rank 0 alone contributes to the all-reduce
            mine = list(range(len(idx))) if self.comm.rank == 0 else []
        else:
            mine = [k for k, i in enumerate(idx) if i // self.per == self.comm.rank]
        if mine and self.depth_sub:
            out[mine] = self.sub.open_batch([idx[k] % self.per for k in mine])
        return out

    def top_paths(self, idx):
        if self.top is None:
            return np.zeros((len(idx), 0, 32), dtype=np.uint8)
        return self.top.open_batch([i // self.per for i in idx])

    def free(self):
        self.sub.free()
        if self.top is not None:
            self.top.free()


def build_tree(backend, comm, values, n, shard_min=1 << 14, width=1):
    """leaf hashing + MerkleTree::new (fri.rs:118-127) over a replicated codeword, sharded by leaf range"""
    G = comm.world
    if G > 1 and width == 1 and n >= shard_min and n % G == 0 and (n // G) >= 2:
        per = n // G
        sub = backend.subtree(values, comm.rank * per, per, 1)
        roots = backend.new_hashes(G)
        comm.all_gather_rows(roots, sub.root)
        return ShardedTree(n, sub, backend.tree_from_hashes(roots), comm)
    return ShardedTree(n, backend.subtree(values, 0, n, width), None, comm)


# ------------------------------------------------------------------------------------------------- FRI

class ShardedFri:
    """Fri (fri.rs:8-311) over G ranks.  Same constructor arguments and panics as Fri::new (fri.rs:30-55)."""

    def __init__(self, backend, omega, offset, domain_length, expansion_factor, num_colinearity_tests, group=None,
                 shard_min=1 << 14):
        self.b, self.comm = backend, Comm(group)
        self.omega, self.offset, self.n = int(omega), int(offset), int(domain_length)
        self.ef, self.nq, self.shard_min = int(expansion_factor), int(num_colinearity_tests), shard_min
        self.rounds = backend.num_rounds(self.n, self.ef, self.nq)      # raises the Fri::new panics
        self.arena = None            # SymmetricArena: enables the fused fold + replicate path (enable_fused_fold)
        self.fused_rounds = 0

    def enable_fused_fold(self):
        """Carve the folded codewords from symmetric memory and let the fold kernel write every replica itself.
        Returns False (and keeps the all-gather path) when symmetric memory cannot be set up."""
        if not (self.comm.nccl and self.comm.world > 1 and hasattr(self.b, "fold_bcast")):
            return False
        try:
            self.arena = SymmetricArena(self.n, self.b.device, self.comm.group)
        except Exception as e:       # no peer access / fabric handles in this environment
            self.arena, self.fused_error = None, repr(e)
            return False
        return True

    def num_rounds(self):
        return self.rounds

    def commit(self, codeword, fs, stream):
        """Fri::commit (fri.rs:105-156): returns (codewords, trees); pushes roots and the last codeword"""
        b, comm = self.b, self.comm
        G, g = comm.world, comm.rank
        om, off = self.omega % P, self.offset % P
        cw, codewords, trees = codeword, [], []
        if self.arena is not None:
            self.arena.reset()
            self.arena.barrier()          # every rank is done with the previous proof's replicas
        for r in range(self.rounds):
            n = cw.numel()
            tree = build_tree(b, comm, cw, n, self.shard_min)
            trees.append(tree)
            root = tree.root_bytes()
            stream.push_merkle_root(root)                    # fri.rs:129-131
            fs.absorb(root)
            if r == self.rounds - 1:                          # fri.rs:133-135
                break
            alpha = fs.challenge()                            # raw u64, fiat_shamir.rs:19-25
            codewords.append(cw)
            h = n // 2
            if self.arena is not None and n >= self.shard_min and h % (4 * G) == 0:
                # fused: the fold kernel stores its slice into every rank's replica (multicast or P2P), then a barrier
                per = h // G
                nxt, peers, mc = self.arena.carve(h)
                b.fold_bcast(cw, n, alpha, off, om, g * per, per, peers, mc)
                self.arena.barrier()
                self.fused_rounds += 1
            elif G > 1 and n >= self.shard_min and h % G == 0:
                nxt = b.new_codeword(h)
                per = h // G
                b.fold_range(cw, n, alpha, off, om, g * per, per, nxt)
                comm.all_gather_inplace(nxt, g * per, per)
            else:
                nxt = b.new_codeword(h)
                b.fold_range(cw, n, alpha, off, om, 0, h, nxt)
            cw = nxt
            om, off = om * om % P, off * off % P              # fri.rs:146-147
        codewords.append(cw)
        stream.push_field_elements(b.download(cw).tolist())    # fri.rs:151
        return codewords, trees

    def prove(self, codeword, transcript=b""):
        """Fri::prove (fri.rs:250-311) + ProofStream::serialize -> (proof bytes, top-level indices)"""
        if codeword.numel() != self.n:
            raise ValueError("initial codeword length does not match domain length")   # fri.rs:256-260
        b, comm = self.b, self.comm
        fs, stream = FiatShamir(b.challenge, transcript), ProofStream()
        codewords, trees = self.commit(codeword, fs, stream)
        seed = b.hash_from_u64(fs.challenge())                                           # fri.rs:272
        size = codewords[1].numel() if len(codewords) > 1 else codewords[0].numel()     # fri.rs:266-270
        top = b.sample_indices(seed, size, codewords[-1].numel(), self.nq)
        # query phase (fri.rs:280-307): collect every (tree, leaf) request, one all-reduce for the sharded lower paths
        reqs, idx = [], list(top)
        for i in range(len(codewords) - 1):
            half = codewords[i].numel() // 2
            idx = [x % half for x in idx]                                                # fri.rs:282-285
            reqs.append((i, list(idx), [x + half for x in idx]))
        local = []
        for i, a, bb in reqs:
            local += [trees[i].local_paths(a), trees[i].local_paths(bb), trees[i + 1].local_paths(a)]
        flat = np.concatenate([x.reshape(-1) for x in local]) if local else np.zeros(0, dtype=np.uint8)
        flat = comm.sum_uint8(flat, getattr(b, "device", "cpu"))
        pos = 0
        for k, (i, a, bb) in enumerate(reqs):
            lower = []
            for x in local[3 * k:3 * k + 3]:
                lower.append(flat[pos:pos + x.size].reshape(x.shape))
                pos += x.size
            upper = [trees[i].top_paths(a), trees[i].top_paths(bb), trees[i + 1].top_paths(a)]
            va, vb = b.gather_values(codewords[i], a), b.gather_values(codewords[i], bb)
            vc = b.gather_values(codewords[i + 1], a)
            for q in range(len(a)):                                                      # fri.rs:229-234
                stream.push_field_elements([va[q], vb[q], vc[q]])
            for q in range(len(a)):                                                      # fri.rs:236-243
                for w in range(3):
                    lo_, up_ = lower[w][q], upper[w][q]
                    stream.push_merkle_path_raw(lo_.shape[0] + up_.shape[0], lo_.tobytes() + up_.tobytes())
        for t in trees:
            t.free()
        return stream.serialize(), top


# ------------------------------------------------------------------------------------------------- LDE + commitment

def lde_commit_sharded(backend, comm, cols_of_group, n_groups, group_width, log_n, log_blowup, offset=3):
    """BASELINE config 4: `n_groups` fixed column groups of `group_width` trace columns; rank g owns groups
    {g, g+G, ...}.  Per owned group: coset LDE of its columns (no communication), one Merkle tree whose leaf i is
    Hash::from_field_elements(row i of the group's LDE) (hash.rs:32-35); all-gather of the group roots; the final
    commitment is MerkleTree::new over the n_groups roots (replicated).

    cols_of_group(k) -> int32 tensor (group_width << log_n), column-major, for an owned group k.
    Returns (commitment bytes, group roots (n_groups, 32) uint8 numpy, {k: LDE tensor} of the owned groups)."""
    G, g = comm.world, comm.rank
    assert n_groups % G == 0, "the fixed group count must be a multiple of the world size"
    N = 1 << (log_n + log_blowup)
    owned = list(range(g, n_groups, G))
    mine = backend.new_hashes(len(owned))
    ldes = {}
    aux = getattr(backend, "aux", None)
    main_stream = torch.cuda.current_stream() if aux is not None else None
    if aux is not None:
        mine.record_stream(aux.stream)
    for j, k in enumerate(owned):
        lde = backend.lde(cols_of_group(k), group_width, log_n, log_blowup, offset)
        if aux is None:
            t = backend.subtree(lde, 0, N, group_width)
            mine[j].copy_(t.root)
            t.free()
        else:
            # the tree of group k on the auxiliary stream, overlapping the LDE of the next group on this one
            aux.stream.wait_stream(main_stream)
            lde.record_stream(aux.stream)
            with torch.cuda.stream(aux.stream):
                t = aux.subtree(lde, 0, N, group_width)
                mine[j].copy_(t.root)
                t.free()
        ldes[k] = lde
    if aux is not None:
        main_stream.wait_stream(aux.stream)
    allr = backend.new_hashes(n_groups).view(G, len(owned), 32)
    comm.all_gather_rows(allr, mine)
    # rank-major (g, j) -> group k = g + j G
    ordered = backend.new_hashes(n_groups)
    for gg in range(G):
        for j in range(len(owned)):
            ordered[gg + j * G].copy_(allr[gg, j])
    top = backend.tree_from_hashes(ordered)
    commitment = top.root_bytes()
    top.free()
    return commitment, ordered.cpu().numpy(), ldes
