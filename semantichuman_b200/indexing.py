"""Host-side index layer: spiral tables, inverse-spiral tables, CSR forms of the D/U sampling matrices.

Everything here runs once at model construction (setup time), through the host entry points of libshb200
(shb_build_inverse_spiral_csr, shb_dense_to_csr, shb_csr_transpose), and leaves int32 tables resident on the device.  Reference contracts: spirals as built by utils_spiral.py:45-95 and cast at main.py:203
((1, V+1, S) int64, -1 == dummy vertex); D/U as padded at main.py:183-205 (dense (1, Vout+1, Vin+1) fp32).
"""
import numpy as np
import torch

from ._capi import check, lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def normalise_spiral(spiral_adj, rows_in=None):
    """(1|B, V+1, S) or (V+1, S), any integer/float dtype, -1 -> rows_in-1 (the Python negative index of
    models.py:42).  Returns a host int32 (V+1, S) array.  Batch-replicated inputs (models.py:122 passes
    S[i].repeat(B,1,1)) are validated once and only the first copy is kept."""
    t = spiral_adj.detach().cpu() if isinstance(spiral_adj, torch.Tensor) else torch.as_tensor(np.asarray(spiral_adj))
    if t.dim() == 3:
        if t.shape[0] > 1 and not bool((t == t[:1]).all()):
            raise ValueError("spiral_adj differs across the batch; SpiralConv expects a batch-replicated table")
        t = t[0]
    if t.dim() != 2:
        raise ValueError("spiral_adj must be (B, V+1, S) or (V+1, S)")
    a = t.numpy().astype(np.int64)
    n = a.shape[0] if rows_in is None else int(rows_in)
    if a.min() < -n or a.max() >= n:
        raise ValueError("spiral index out of range")
    return _i32(np.where(a < 0, a + n, a))


def build_inverse_spiral_csr(table, rows_in):
    """SURVEY 8(a-8) inverse-spiral CSR: (rowptr (rows_in+1), slots (rows_out*S)) host int32."""
    table = _i32(table)
    rows_out, S = table.shape
    rowptr = np.empty(rows_in + 1, np.int32)
    slots = np.empty(rows_out * S, np.int32)
    check(lib.shb_build_inverse_spiral_csr(table.ctypes.data, rows_out, S, rows_in, rowptr.ctypes.data,
                                           slots.ctypes.data), "shb_build_inverse_spiral_csr")
    return rowptr, slots


def build_conv_groups(ptr, ent, rows_dst, rows_src, R, SPS):
    """Shared-source group program of a SpiralConv pass (shb_build_conv_groups): host arrays (gptr, recs, gdst, gmask)."""
    import ctypes

    ptr, ent = _i32(ptr), _i32(ent)
    ng, nr = ctypes.c_int32(0), ctypes.c_int32(0)
    args = (ptr.ctypes.data, ent.ctypes.data, int(rows_dst), int(rows_src), int(R), int(SPS), ctypes.addressof(ng),
            ctypes.addressof(nr))
    check(lib.shb_build_conv_groups(*args, None, None, None, None), "shb_build_conv_groups")
    gptr = np.empty(ng.value + 1, np.int32)
    recs = np.empty(nr.value * 48, np.int32)
    gdst = np.empty(ng.value * R, np.int32)
    gmask = np.empty(ng.value, np.uint32)
    check(lib.shb_build_conv_groups(*args, gptr.ctypes.data, recs.ctypes.data, gdst.ctypes.data, gmask.ctypes.data),
          "shb_build_conv_groups")
    return gptr, recs, gdst, gmask


class GroupProgram:
    """Device-resident group program of one pass of one geometry (see build_conv_groups)."""

    __slots__ = ("gptr", "recs", "gdst", "gmask", "n_groups", "n_records", "R", "SPS", "n_loads")

    def __init__(self, ptr, ent, rows_dst, rows_src, R, SPS, device):
        gptr, recs, gdst, gmask = build_conv_groups(ptr, ent, rows_dst, rows_src, R, SPS)
        self.n_groups, self.n_records, self.R, self.SPS = len(gmask), len(recs) // 48, int(R), int(SPS)
        self.n_loads = int((recs.reshape(-1, 48)[:, 0] & 15).sum())  # slab loads per batch chunk (vs. one per entry ungrouped)
        self.gptr = torch.from_numpy(gptr).to(device)
        self.recs = torch.from_numpy(recs).to(device)
        self.gdst = torch.from_numpy(gdst).to(device)
        self.gmask = torch.from_numpy(gmask.view(np.int32)).to(device)


def dense_to_csr(dense):
    """fp32 dense (rows, cols) -> (rowptr, colidx, vals) host arrays, exact zeros dropped."""
    d = np.ascontiguousarray(dense, dtype=np.float32)
    rows, cols = d.shape
    nnz = np.zeros(1, np.int64)
    check(lib.shb_dense_to_csr(d.ctypes.data, rows, cols, None, None, None, 0, nnz.ctypes.data), "shb_dense_to_csr")
    rowptr = np.empty(rows + 1, np.int32)
    colidx = np.empty(max(int(nnz[0]), 1), np.int32)
    vals = np.empty(max(int(nnz[0]), 1), np.float32)
    check(lib.shb_dense_to_csr(d.ctypes.data, rows, cols, rowptr.ctypes.data, colidx.ctypes.data, vals.ctypes.data,
                               int(nnz[0]), nnz.ctypes.data), "shb_dense_to_csr")
    return rowptr, colidx[: int(nnz[0])], vals[: int(nnz[0])]


def csr_transpose(rowptr, colidx, vals, rows, cols):
    rowptr, colidx = _i32(rowptr), _i32(colidx)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    t_rowptr = np.empty(cols + 1, np.int32)
    t_colidx = np.empty(max(len(colidx), 1), np.int32)
    t_vals = np.empty(max(len(vals), 1), np.float32)
    if len(colidx) == 0:
        t_rowptr[:] = 0
        return t_rowptr, t_colidx[:0], t_vals[:0]
    check(lib.shb_csr_transpose(rowptr.ctypes.data, colidx.ctypes.data, vals.ctypes.data, rows, cols,
                                t_rowptr.ctypes.data, t_colidx.ctypes.data, t_vals.ctypes.data), "shb_csr_transpose")
    return t_rowptr, t_colidx, t_vals


def locality_order(table):
    """Vertex order that makes spiral neighbourhoods index-local: reverse Cuthill-McKee on the graph whose edges are the
    (vertex, spiral entry) pairs of `table` ((V+1, S) normalised host table; the dummy row/entry V is left out).
    Returns `perm` (V,) -- new position i holds old vertex perm[i].  A 128-row tile of the conv then gathers a few hundred
    distinct rows instead of ~1300 when the mesh's own numbering is arbitrary (measured: forward -19 %, wgrad -49 % at
    level 0 of the 6890-vertex template)."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee

    table = np.asarray(table)
    V = table.shape[0] - 1
    rows = np.repeat(np.arange(V), table.shape[1])
    cols = table[:V].ravel()
    keep = cols < V
    a = sp.coo_matrix((np.ones(int(keep.sum()), np.int8), (rows[keep], cols[keep])), shape=(V, V)).tocsr()
    a = ((a + a.T) > 0).astype(np.int8).tocsr()
    return np.asarray(reverse_cuthill_mckee(a, symmetric_mode=True), dtype=np.int64)


class SpiralGeometry:
    """Entry lists of one SpiralConv call shape, resident on the device (the kernels' view of a spiral table).

    table    (rows_out, S) int32: source row of x for output row j, slot s (already -1 -> rows_in-1); rows_out < rows_in
             when the conv is fused with a selection down-pool (output rows = kept vertices + dummy).
    forward : for every output row j the entries (table[j,s] << 5 | s); entries reading the source's dummy row are dropped
              when that row is known to be zero (`src_dummy_zero`: its producer masked it).
    backward: for every source row u the entries (j << 5 | s) with table[j,s] == u, ascending in (j, s) -- the inverse-
              spiral CSR of SURVEY 8 a-8 (shb_build_inverse_spiral_csr), i.e. the fixed summation order of the input
              gradient; rows whose gz is zero by the mask (`zero_last_row`) are dropped, and the dummy source row gets no
              entries unless its gradient is wanted (`dummy_row_grad`).
    """

    def __init__(self, table, rows_in, device, zero_last_row=True, dummy_row_grad=True, src_dummy_zero=False):
        table = _i32(table)
        self.rows_out, self.S = int(table.shape[0]), int(table.shape[1])
        self.rows_in = int(rows_in)
        if self.S > 32:
            raise ValueError("the kernels support spiral lengths up to 32")
        if table.min() < 0 or table.max() >= rows_in:
            raise ValueError("spiral index out of range")
        self.zero_last_row, self.dummy_row_grad, self.src_dummy_zero = bool(zero_last_row), bool(dummy_row_grad), bool(src_dummy_zero)
        self.table_host = table
        self.device = torch.device(device)
        S, dummy = self.S, self.rows_in - 1
        slots = np.tile(np.arange(S, dtype=np.int64), self.rows_out)
        flat = table.reshape(-1).astype(np.int64)
        # forward lists
        keep = np.ones(flat.shape, bool) if not self.src_dummy_zero else flat != dummy
        counts = keep.reshape(self.rows_out, S).sum(1)
        self._host_f = (_i32(np.concatenate([[0], np.cumsum(counts)])), _i32(((flat << 5) | slots)[keep]))
        self.ptr_f = self._dev(self._host_f[0])
        self.ent_f = self._dev(self._host_f[1])
        # backward lists: the canonical inverse CSR (flat positions j*S+s, ascending per source row), then the dead
        # entries taken out
        rowptr, pos = build_inverse_spiral_csr(table, self.rows_in)
        pos = pos.astype(np.int64)
        src_of = np.repeat(np.arange(self.rows_in, dtype=np.int64), np.diff(rowptr))
        jj, ss = pos // S, pos % S
        live = np.ones(pos.shape, bool)
        if self.zero_last_row:
            live &= jj != self.rows_out - 1
        if not self.dummy_row_grad:
            live &= src_of != dummy
        # A live dummy source row (the FC row under the first decoder conv) is referenced by every padded spiral entry:
        # hundreds of entries in ONE list, i.e. one tile that a single CTA would grind through.  Its list is cut into
        # sub-lists that run as ordinary tiles into a scratch tensor (one partial row each); a Pool row then adds the
        # partials in order and applies the producer's activation derivative.
        self.dummy_split = None
        to_dummy = live & (src_of == dummy)
        n_dummy = int(to_dummy.sum())
        if self.dummy_row_grad and n_dummy > 4 * S:
            live &= src_of != dummy
            ents = (jj[to_dummy] << 5) | ss[to_dummy]
            T = (n_dummy + 47) // 48
            bounds = (np.arange(T + 1, dtype=np.int64) * n_dummy) // T
            self.dummy_split = (T, self._dev(bounds), self._dev(ents), self._dev(np.array([0, T])), self._dev(np.arange(T)),
                                torch.ones(T, dtype=torch.float32, device=self.device))
        counts_b = np.bincount(src_of[live], minlength=self.rows_in)
        self._host_b = (_i32(np.concatenate([[0], np.cumsum(counts_b)])), _i32(((jj << 5) | ss)[live]))
        self.ptr_b = self._dev(self._host_b[0])
        self.ent_b = self._dev(self._host_b[1])
        self._programs = {}
        self.table = torch.from_numpy(table).to(self.device)
        self.n_fwd_entries, self.n_bwd_entries = int(counts.sum()), int(counts_b.sum())

    def group_program(self, backward, R, SPS):
        """Shared-source group program of the forward (destinations = output rows) or input-gradient (destinations = source
        rows) pass for group size R and SPS slabs per record; built on first use, cached."""
        key = (bool(backward), int(R), int(SPS))
        prog = self._programs.get(key)
        if prog is None:
            if len(self._programs) > 8:
                self._programs.clear()
            ptr, ent = self._host_b if backward else self._host_f
            rows_dst, rows_src = (self.rows_in, self.rows_out) if backward else (self.rows_out, self.rows_in)
            prog = GroupProgram(ptr, ent, rows_dst, rows_src, R, SPS, self.device)
            self._programs[key] = prog
        return prog

    def _dev(self, a):
        a = np.ascontiguousarray(a, dtype=np.int64)
        if a.size and a.max() >= 2 ** 31:
            raise ValueError("table too large for 32-bit entry lists")
        if a.size == 0:
            a = np.zeros(1, np.int64)
        return torch.from_numpy(a.astype(np.int32)).to(self.device)

    @classmethod
    def from_spiral(cls, spiral_adj, device, **kw):
        table = normalise_spiral(spiral_adj)
        return cls(table, table.shape[0], device, **kw)

    def _flags(self, kw):
        f = dict(zero_last_row=self.zero_last_row, dummy_row_grad=self.dummy_row_grad, src_dummy_zero=self.src_dummy_zero)
        f.update({k: bool(v) for k, v in kw.items()})
        return f

    def restricted(self, out_rows, **kw):
        """Geometry that evaluates only `out_rows` (source-vertex ids, dummy last)."""
        out_rows = np.asarray(out_rows, dtype=np.int64)
        return SpiralGeometry(self.table_host[out_rows], self.rows_in, self.device, **self._flags(kw))

    def with_flags(self, **kw):
        """Same table, different promises about the dummy rows (the entry lists depend on them)."""
        f = self._flags(kw)
        if all(getattr(self, k) == v for k, v in f.items()):
            return self
        return SpiralGeometry(self.table_host, self.rows_in, self.device, **f)


class PoolMatrix:
    """A D or U sampling matrix in CSR form (forward) plus the CSR of its transpose (backward), on the device."""

    def __init__(self, rowptr, colidx, vals, rows, cols, device):
        self.rows_out, self.rows_in = int(rows), int(cols)
        rowptr, colidx = _i32(rowptr), _i32(colidx)
        vals = np.ascontiguousarray(vals, dtype=np.float32)
        self.nnz = int(len(colidx))
        self._host_csr = (rowptr, np.ascontiguousarray(colidx), vals)
        counts = np.diff(rowptr)
        # structural fact the fused conv+down-pool relies on (mesh_sampling.py:214-227): one 1.0 per row
        self.is_selection = bool(self.nnz == rows and (counts == 1).all() and (vals == 1.0).all())
        self.selection_cols = colidx.copy() if self.is_selection else None
        # last row == e_dummy (main.py:190-191): the pooled tensor's dummy row is the input's dummy row, nothing else
        self.dummy_preserving = bool(rows > 0 and counts[-1] == 1 and colidx[rowptr[rows - 1]] == cols - 1
                                     and vals[rowptr[rows - 1]] == 1.0)
        t_rowptr, t_colidx, t_vals = csr_transpose(rowptr, colidx, vals, rows, cols)
        self.device = torch.device(device)
        dev = self.device
        self.rowptr = torch.from_numpy(rowptr).to(dev)
        self.colidx = torch.from_numpy(np.ascontiguousarray(colidx)).to(dev)
        self.vals = torch.from_numpy(vals).to(dev)
        self.t_rowptr = torch.from_numpy(t_rowptr).to(dev)
        self.t_colidx = torch.from_numpy(np.ascontiguousarray(t_colidx)).to(dev)
        self.t_vals = torch.from_numpy(np.ascontiguousarray(t_vals)).to(dev)

    def permuted(self, new_rows, new_cols):
        """P' = P[new_rows][:, new_cols]: row i of the result is row new_rows[i] of P, column c is column new_cols[c]
        (both full permutations, dummy index included).  Used by the locality re-ordering of the model trunks."""
        rowptr, colidx, vals = self._host_csr
        new_rows = np.asarray(new_rows, dtype=np.int64)
        col_pos = np.empty(self.rows_in, np.int64)
        col_pos[np.asarray(new_cols, dtype=np.int64)] = np.arange(self.rows_in)
        counts = np.diff(rowptr)[new_rows]
        out_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        out_col = np.empty(self.nnz, np.int32)
        out_val = np.empty(self.nnz, np.float32)
        for i, r in enumerate(new_rows):  # setup time; rows are short (1-3 entries)
            a, b = rowptr[r], rowptr[r + 1]
            c = col_pos[colidx[a:b]]
            o = np.argsort(c, kind="stable")
            out_col[out_ptr[i]:out_ptr[i + 1]] = c[o]
            out_val[out_ptr[i]:out_ptr[i + 1]] = vals[a:b][o]
        return PoolMatrix(out_ptr, out_col, out_val, self.rows_out, self.rows_in, self.device)

    @classmethod
    def from_permutation(cls, perm_full, device):
        """Row gather y[:, i] = x[:, perm_full[i]] as a selection-type PoolMatrix (its transpose is the inverse gather)."""
        n = len(perm_full)
        return cls(np.arange(n + 1, dtype=np.int32), _i32(perm_full), np.ones(n, np.float32), n, n, device)

    @classmethod
    def from_dense(cls, dense, device=None):
        """dense: (1, rows, cols) or (rows, cols) tensor/array exactly as main.py:183-205 builds it."""
        if isinstance(dense, torch.Tensor):
            device = dense.device if device is None else device
            d = dense.detach().to("cpu", torch.float32).numpy()
        else:
            d = np.asarray(dense, dtype=np.float32)
        if d.ndim == 3:
            if d.shape[0] != 1:
                raise ValueError("sampling matrix must have a leading dimension of 1")
            d = d[0]
        rowptr, colidx, vals = dense_to_csr(d)
        return cls(rowptr, colidx, vals, d.shape[0], d.shape[1], device)

    @classmethod
    def from_scipy_padded(cls, m, device):
        """scipy sparse (Vout, Vin) un-padded matrix -> padded CSR with the dummy->dummy corner (main.py:190-191),
        explicit zeros dropped (the dense path cannot see them either)."""
        import scipy.sparse as sp

        m = sp.csr_matrix(m).astype(np.float32)
        m.eliminate_zeros()
        m.sort_indices()
        rows, cols = m.shape
        rowptr = np.concatenate([m.indptr, [m.indptr[-1] + 1]]).astype(np.int32)
        colidx = np.concatenate([m.indices, [cols]]).astype(np.int32)
        vals = np.concatenate([m.data, [1.0]]).astype(np.float32)
        return cls(rowptr, colidx, vals, rows + 1, cols + 1, device)
