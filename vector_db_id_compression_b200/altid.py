"""Host-side mirror of the reference's `altid` module (alt-graph-index/altid.swig, altid_impl.{h,cpp}):
compressed NSG adjacency with the reference's class names. `FinalNSGGraph` stands in for
faiss::nsg::Graph<int32_t> (data N x K int32, rows terminated by the first -1)."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import capi
from .custom_invlists import default_context


class FinalNSGGraph:
    def __init__(self, data: np.ndarray):
        self.data = np.ascontiguousarray(data, dtype=np.int32)
        self.N, self.K = self.data.shape

    def get_neighbors(self, i: int):
        row = self.data[i]
        stop = np.nonzero(row == -1)[0]
        n = int(stop[0]) if stop.size else self.K
        return n, row[:n]


class _CompressedGraph:
    def __init__(self, graph: FinalNSGGraph, ctx: Optional[capi.Context]):
        self.ctx = ctx or default_context()
        self.N, self.K = graph.N, graph.K
        self.data = None  # altid_impl.cpp:38,89,150: the compressed graph owns no raw rows
        self.compressed_ids_size_in_bytes = 0
        self.overhead_in_bytes = 0

    def get_neighbors_batch(self, rows):
        """-> (neighbors [m, K] padded with -1, counts [m]) for many rows in one GPU call."""
        return self.blob.decode_rows(np.asarray(rows, dtype=np.int32))


class EliasFanoNSGGraph(_CompressedGraph):
    """altid_impl.h:42-50, .cpp:53-101."""

    def __init__(self, graph: FinalNSGGraph, ctx: Optional[capi.Context] = None):
        super().__init__(graph, ctx)
        # size of each friend list + max id value, altid_impl.cpp:56-57
        # two `size_t += double` statements: each truncates (N = 100: 87 + 87, not 175)
        self.overhead_in_bytes = 2 * int(self.N * math.ceil(math.log2(self.N)) / 8.0) if self.N > 1 else 0
        self.blob = self.ctx.ef_encode_rows(graph.data)
        self.compressed_ids_size_in_bytes = self.blob.bits_total // 8  # :86-88

    def get_neighbors(self, i: int):
        nb, cnt = self.blob.decode_rows(np.asarray([i], dtype=np.int32))
        n = int(cnt[0])
        return n, nb[0, :n]  # returns ef->num_elements, altid_impl.cpp:100


class ROCNSGGraph(_CompressedGraph):
    """altid_impl.h:53-67, .cpp:103-165."""

    def __init__(self, graph: FinalNSGGraph, ctx: Optional[capi.Context] = None):
        super().__init__(graph, ctx)
        self.overhead_in_bytes = int(self.N * math.ceil(math.log2(self.N)) / 8.0) if self.N > 1 else 0  # :106
        self.blob = self.ctx.roc_encode_rows(graph.data)
        ex = self.blob.export()
        self.num_outgoing_edges = ex["unit_n"].astype(np.uint32)
        self.id_symbol_precision = ex["precision"].astype(np.uint64)
        self.compressed_ids_size_in_bytes = self.blob.ans_bytes  # :148

    def get_neighbors(self, i: int):
        nb, cnt = self.blob.decode_rows(np.asarray([i], dtype=np.int32))
        n = int(cnt[0])
        return self.K, nb[0, :n]  # the reference returns K, not n (altid_impl.cpp:164)


class CompactBitNSGGraph(_CompressedGraph):
    """altid_impl.h:29-39, .cpp:20-51: every edge in ceil(log2(N+1)) bits, N marks the end of a row."""

    def __init__(self, graph: FinalNSGGraph, ctx: Optional[capi.Context] = None):
        super().__init__(graph, ctx)
        self.bits = 0
        while (1 << self.bits) < self.N + 1:
            self.bits += 1
        self.stride = (self.K * self.bits + 7) // 8
        vals = graph.data.astype(np.int64)
        marker = vals == -1
        ended = np.cumsum(marker, axis=1)       # >= 1 from the first -1 of a row on
        vals[ended >= 1] = 0                    # nothing is written behind the end marker: the bytes stay zero
        vals[marker & (ended == 1)] = self.N    # writer.write(N, bits); break;  (altid_impl.cpp:30-33)
        # one packed string per row, each padded to `stride` bytes
        per_row_bits = self.stride * 8
        flat = self.ctx.bits_pack(vals.ravel().astype(np.uint64), self.bits) if per_row_bits == self.K * self.bits else None
        if flat is not None:
            self.compressed_data = np.ascontiguousarray(flat).reshape(self.N, self.stride)
        else:
            self.compressed_data = np.stack([self.ctx.bits_pack(vals[i].astype(np.uint64), self.bits, self.stride)
                                             for i in range(self.N)])
        self.compressed_ids_size_in_bytes = self.N * self.stride

    def get_neighbors(self, i: int):
        row = self.ctx.bits_unpack(self.compressed_data[i], self.K, self.bits).astype(np.int64)
        stop = np.nonzero(row == self.N)[0]
        n = int(stop[0]) if stop.size else self.K
        return n, row[:n].astype(np.int32)
