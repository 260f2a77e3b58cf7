"""Test harness: the call sequences of the distributed set-up entry points (partition,
separate_local_nonlocal, index map) over a backend of tests/helpers.py -- the oracle
(`orc_*`), the host-compiled kernel source, or the CUDA library (`b200_*`).  numpy in, numpy
out; host out-parameters are helpers.OutInt / OutI64 boxes."""
import numpy as np

from tests import helpers as H

NP = {"i32": np.int32, "i64": np.int64, "f64": np.float64, "f32": np.float32}


def _pad(a):
    """backends dislike zero-length buffers: keep one spare element"""
    return a if a.size else np.zeros(1, a.dtype)


class Partition:
    def __init__(self, lt, gt, range_bounds, part_ids, num_parts):
        self.lt, self.gt = lt, gt
        self.range_bounds = np.ascontiguousarray(range_bounds, NP[gt])
        self.part_ids = np.ascontiguousarray(part_ids, np.int32)
        self.num_parts = int(num_parts)
        self.num_ranges = len(self.part_ids)
        self.size = int(self.range_bounds[-1])
        self.starting_indices = self.part_sizes = None
        self.num_empty_parts = None

    def finalize(self, be):
        """Partition::finalize_construction (core/distributed/partition.cpp:122-137)"""
        start = np.zeros(max(self.num_ranges, 1), NP[self.lt])
        sizes = np.zeros(max(self.num_parts, 1), NP[self.lt])
        empty = H.OutInt()
        be("partition_build_starting_indices_%s_%s" % (self.lt, self.gt), self.num_ranges, self.num_parts,
           _pad(self.range_bounds), _pad(self.part_ids), start, sizes, empty)
        self.starting_indices = start[:self.num_ranges]
        self.part_sizes = sizes[:self.num_parts]
        self.num_empty_parts = empty.value
        ordered = H.OutInt()
        be("partition_has_ordered_parts", self.num_ranges, _pad(self.part_ids), ordered)
        self.ordered = bool(ordered.value)
        return self

    def as_dict(self):
        return dict(size=self.size, num_ranges=self.num_ranges, num_parts=self.num_parts,
                    num_empty_parts=self.num_empty_parts, ordered=self.ordered,
                    range_bounds=self.range_bounds, part_ids=self.part_ids,
                    starting_indices=self.starting_indices, part_sizes=self.part_sizes)


def partition_from_mapping(be, mapping, num_parts, lt="i32", gt="i64"):
    mapping = np.ascontiguousarray(mapping, np.int32)
    n = len(mapping)
    nr = H.OutI64()
    be("partition_count_ranges", n, _pad(mapping), nr)
    bounds = np.zeros(nr.value + 1, NP[gt])
    ids = np.zeros(max(nr.value, 1), np.int32)
    be("partition_build_from_mapping_" + gt, n, _pad(mapping), bounds, ids)
    return Partition(lt, gt, bounds, ids[:nr.value], num_parts).finalize(be)


def partition_from_contiguous(be, ranges, part_ids=None, lt="i32", gt="i64"):
    ranges = np.ascontiguousarray(ranges, NP[gt])
    n = len(ranges) - 1
    pm = None if part_ids is None else np.ascontiguousarray(part_ids, np.int32)
    bounds = np.zeros(n + 1, NP[gt])
    ids = np.zeros(max(n, 1), np.int32)
    be("partition_build_from_contiguous_" + gt, n, ranges, 0 if pm is None else _pad(pm), bounds, ids)
    return Partition(lt, gt, bounds, ids[:n], n).finalize(be)


def partition_uniform(be, num_parts, global_size, lt="i32", gt="i64"):
    ranges = np.zeros(num_parts + 1, NP[gt])
    be("partition_build_ranges_from_global_size_" + gt, num_parts, global_size, ranges)
    if num_parts == 0:
        ranges[:] = 0
    return partition_from_contiguous(be, ranges, None, lt, gt)


def separate(be, rp, cp, rows, cols, vals, local_part, vt="f64"):
    """-> dict(local=(rows, cols, vals), non_local=(rows, global cols, vals), kept=(...), cls, ranks)"""
    lt, gt = rp.lt, rp.gt
    rows = np.ascontiguousarray(rows, NP[gt])
    cols = np.ascontiguousarray(cols, NP[gt])
    vals = np.ascontiguousarray(vals, NP[vt])
    nnz = len(rows)
    cls = np.zeros(max(nnz, 1), np.uint8)
    lrank = np.zeros(nnz + 1, np.int64)
    nrank = np.zeros(nnz + 1, np.int64)
    nl, nn = H.OutI64(), H.OutI64()
    be("dist_classify_entries_" + gt, nnz, _pad(rows), _pad(cols), rp.num_ranges, rp.range_bounds,
       _pad(rp.part_ids), cp.num_ranges, cp.range_bounds, _pad(cp.part_ids), local_part, cls, lrank, nrank,
       nl, nn)
    a, b = nl.value, nn.value

    def z(n, t):
        return np.zeros(max(n, 1), NP[t])

    lr, lc, lv = z(a, lt), z(a, lt), z(a, vt)
    nr, nc, nv = z(b, lt), z(b, gt), z(b, vt)
    be("dist_separate_fill_%s_%s_%s" % (vt, lt, gt), nnz, _pad(rows), _pad(cols), _pad(vals), rp.num_ranges,
       rp.range_bounds, _pad(rp.starting_indices), cp.num_ranges, cp.range_bounds, _pad(cp.starting_indices),
       cls, lrank, nrank, lr, lc, lv, nr, nc, nv)
    kr, kc, kv = z(a + b, lt), z(a + b, gt), z(a + b, vt)
    be("dist_kept_fill_%s_%s_%s" % (vt, lt, gt), nnz, _pad(rows), _pad(cols), _pad(vals), rp.num_ranges,
       rp.range_bounds, _pad(rp.starting_indices), cls, lrank, nrank, kr, kc, kv)
    return dict(local=(lr[:a], lc[:a], lv[:a]), non_local=(nr[:b], nc[:b], nv[:b]),
                kept=(kr[:a + b], kc[:a + b], kv[:a + b]), cls=cls[:nnz], local_rank=lrank,
                non_local_rank=nrank)


class IndexMap:
    """index_map(exec, partition, rank, recv_connections) (core/distributed/index_map.cpp)"""

    def __init__(self, be, part, rank, conns, skip_part=-1):
        self.part, self.rank = part, rank
        lt, gt = part.lt, part.gt
        conns = np.ascontiguousarray(conns, NP[gt])
        words = (part.size + 31) // 32
        # the bit pattern of the uint32 words, carried as int32 (torch stages it)
        self.bitmap = np.zeros(words + 1, np.int32)
        be("index_map_mark_" + gt, part.size, part.num_ranges, part.range_bounds, _pad(part.part_ids),
           skip_part, len(conns), _pad(conns), self.bitmap)
        self.word_rank = np.zeros(words + 1, np.int64)
        self.range_offsets = np.zeros(max(part.num_ranges, 1), np.int64)
        self.remote_sizes = np.zeros(max(part.num_parts, 1), np.int64)
        nrem = H.OutI64()
        be("index_map_rank_" + gt, part.size, part.num_ranges, part.num_parts, part.range_bounds,
           _pad(part.part_ids), self.bitmap, self.word_rank, self.range_offsets, self.remote_sizes, nrem)
        self.num_remote = nrem.value
        self.remote_sizes = self.remote_sizes[:part.num_parts]
        n = self.num_remote
        self.remote_global = np.zeros(max(n, 1), NP[gt])
        self.remote_local = np.zeros(max(n, 1), NP[lt])
        self.remote_part_ids = np.zeros(max(n, 1), np.int32)
        be("index_map_fill_%s_%s" % (lt, gt), part.size, part.num_ranges, part.range_bounds,
           _pad(part.part_ids), _pad(part.starting_indices), self.bitmap, self.word_rank, self.range_offsets,
           self.remote_global, self.remote_local, self.remote_part_ids)
        self.remote_global = self.remote_global[:n]
        self.remote_local = self.remote_local[:n]
        self.remote_part_ids = self.remote_part_ids[:n]

    def target_ids(self):
        """the reference's compressed view: parts with at least one remote index, their sizes"""
        ids = np.nonzero(self.remote_sizes)[0].astype(np.int32)
        return ids, self.remote_sizes[ids]

    def map_to_local(self, be, global_ids, index_space):
        part = self.part
        lt, gt = part.lt, part.gt
        g = np.ascontiguousarray(global_ids, NP[gt])
        out = np.zeros(max(len(g), 1), NP[lt])
        local_size = int(part.part_sizes[self.rank]) if part.num_parts else 0
        be("index_map_map_to_local_%s_%s" % (lt, gt), part.size, part.num_ranges, part.range_bounds,
           _pad(part.part_ids), _pad(part.starting_indices), self.bitmap, self.word_rank, self.range_offsets,
           self.rank, local_size, index_space, len(g), _pad(g), out)
        return out[:len(g)]


def vector_build_local(be, part, rows, cols, vals, ncols, local_part, vt="f64"):
    """distributed_vector::build_local -> the n_local x ncols block of part local_part"""
    lt, gt = part.lt, part.gt
    rows = np.ascontiguousarray(rows, NP[gt])
    cols = np.ascontiguousarray(cols, NP[gt])
    vals = np.ascontiguousarray(vals, NP[vt])
    n_local = int(part.part_sizes[local_part])
    local = np.zeros(max(n_local * ncols, 1), NP[vt])
    be("dist_vector_build_local_%s_%s_%s" % (vt, lt, gt), len(rows), _pad(rows), _pad(cols), _pad(vals),
       part.num_ranges, part.range_bounds, _pad(part.part_ids), _pad(part.starting_indices), local_part, local,
       ncols)
    return local[:n_local * ncols].reshape(n_local, ncols)
