"""Drop-in replacements for the reference's ``models.py`` classes, running on libshb200 kernels.

Same constructor and forward signatures, same ``state_dict`` keys (``conv.{j}.conv.weight`` ...,
``fc_latent_enc.*``, ``fc_latent_dec.*``, ``fc_latent_enc_list.{k}.*``, ``fc_latent_dec_list.{k}.*``,
``kps_enc_list.{k}.*``; models.py:81-86,113,198-204,230), so the reference's training loop (train_funcs.py), entry
script (main.py:241-259) and editing script (demo.py / utils_SH.py:358-376) can construct and call them unchanged
and load the reference's checkpoints.

What is different underneath
  * Inside the trunks activations live batch-innermost in 128-sample slabs (slab.py): SpiralConv is ONE fused kernel that
    moves the S neighbour slabs of an output vertex by TMA straight into tensor-core operands (+bias +activation
    +dummy-row mask); the gathered (B*(V+1), S*Cin) matrix of models.py:42 is never materialised.
  * The dense D/U ``torch.matmul`` of models.py:127,148 becomes a CSR sum of whole slabs; D matrices that are pure row
    selections (mesh_sampling.py:214-227) are folded into the preceding conv, which then only evaluates the kept
    vertices (half the work).
  * Index tables are built once per level at construction; ``S[i].repeat(B,1,1)`` (models.py:122) is never formed.
  * No CPU path: CPU tensors raise.
"""
import copy

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functions as fn
from . import slab
from ._capi import ACT_ENUM
from .indexing import PoolMatrix, SpiralGeometry, locality_order, normalise_spiral

# models.py:169,285 read cfg.CONSTANTS.newskl_list; this is its default (configure/cfgs.py:21-23).  A caller that
# merges a yaml with a longer list (traincfg.yaml:55) can assign `model.newskl_list`.
DEFAULT_NEWSKL_LIST = [[0, 1], [0, 2], [0, 6], [1, 4], [2, 5], [6, 9], [4, 7], [5, 8], [9, 12], [9, 16], [9, 17],
                       [7, 10], [8, 11], [12, 15], [16, 18], [17, 19], [18, 20], [19, 21], [20, 22], [21, 23],
                       [20, 24], [21, 25], [20, 26], [21, 27], [15, 28], [15, 29], [15, 30]]


class SpiralConv(nn.Module):
    """models.py:10-53.  ``forward(x, spiral_adj)`` accepts the reference's (B, V+1, S) int64 tensor (validated
    and cached per table) or, on the fast path, a prebuilt :class:`SpiralGeometry`."""

    def __init__(self, in_c, spiral_size, out_c, activation='elu', bias=True, device=None):
        super().__init__()
        if activation not in ACT_ENUM:
            raise NotImplementedError()  # models.py:31-32
        self.in_c = in_c
        self.out_c = out_c
        self.spiral_size = spiral_size
        self.device = device
        self.activation_name = activation
        self.conv = nn.Linear(in_c * spiral_size, out_c, bias=bias)
        self._geom_cache = []  # [(first batch copy of the table on device, SpiralGeometry)]

    def _geometry_for(self, spiral_adj):
        first = spiral_adj[0] if spiral_adj.dim() == 3 else spiral_adj
        for src, geom in self._geom_cache:
            if src.shape == first.shape and src.device == first.device and torch.equal(src, first):
                return geom
        geom = SpiralGeometry.from_spiral(spiral_adj, spiral_adj.device)  # validates batch invariance once
        self._geom_cache.append((first.clone(), geom))
        return geom

    def forward(self, x, spiral_adj, compute_dtype=None):
        geom = spiral_adj if isinstance(spiral_adj, SpiralGeometry) else self._geometry_for(spiral_adj)
        return fn.spiral_conv(x, self.conv.weight, self.conv.bias, geom, self.activation_name, compute_dtype)


class Pool(nn.Module):
    """The reference has no Pool class: this names the ``torch.matmul(D[i], x)`` / ``torch.matmul(U[i], x)`` call
    sites (models.py:127,148,250,276).  Accepts the dense padded (1, Vout+1, Vin+1) tensor of main.py:183-205
    (converted to CSR once), a scipy sparse matrix (un-padded), or a :class:`PoolMatrix`."""

    def __init__(self, matrix, device=None):
        super().__init__()
        self.pm = as_pool_matrix(matrix, device)

    def forward(self, x):
        return fn.pool(x, self.pm)


def as_pool_matrix(m, device=None):
    if isinstance(m, PoolMatrix):
        return m
    if isinstance(m, torch.Tensor) or isinstance(m, np.ndarray):
        return PoolMatrix.from_dense(m, device)
    return PoolMatrix.from_scipy_padded(m, device)


def _conv_stacks(filters_enc, filters_dec, spiral_sizes, activation, device):
    """Layer construction order of models.py:69-113 (also :186-230).  Returns (enc_convs, enc_levels, dec_convs,
    dec_levels, last_enc_channels)."""
    n_levels = len(spiral_sizes) - 1
    enc, enc_lvl, dec, dec_lvl = [], [], [], []
    c = filters_enc[0][0]
    for i in range(n_levels):
        for width in ([filters_enc[1][i]] if filters_enc[1][i] else []) + [filters_enc[0][i + 1]]:
            enc.append(SpiralConv(c, spiral_sizes[i], width, activation=activation, device=device).to(device))
            enc_lvl.append(i)
            c = width
    enc_out = c
    c = filters_dec[0][0]
    for i in range(n_levels):
        lvl = n_levels - 1 - i
        widths = [filters_dec[0][i + 1]] + ([filters_dec[1][i + 1]] if filters_dec[1][i + 1] else [])
        for q, width in enumerate(widths):
            last = (i == n_levels - 1) and (q == len(widths) - 1)
            dec.append(SpiralConv(c, spiral_sizes[lvl], width, activation='identity' if last else activation,
                                  device=device).to(device))
            dec_lvl.append(lvl)
            c = width
    return enc, enc_lvl, dec, dec_lvl, enc_out


class _SpiralTrunk(nn.Module):
    """Conv/pool stacks shared by both autoencoders (models.py:121-127,147-153 == :244-250,275-281)."""

    def _init_trunk(self, filters_enc, filters_dec, sizes, spiral_sizes, spirals, D, U, device, activation, fuse_pool,
                    make_heads, reorder=True):
        """`make_heads(enc_out_channels)` registers the latent layers; it runs between the registration of
        ``conv`` and ``dconv`` so that parameter order (hence optimizer-state order in the reference's checkpoints,
        main.py:288) is the reference's: conv, fc..., dconv (models.py:81-86,113)."""
        self.sizes = sizes
        self.spirals = spirals
        self.filters_enc = filters_enc
        self.filters_dec = filters_dec
        self.spiral_sizes = spiral_sizes
        self.D = D
        self.U = U
        self.device = device
        self.activation = activation
        enc, self._enc_lvl, dec, self._dec_lvl, self._enc_out = _conv_stacks(filters_enc, filters_dec, spiral_sizes,
                                                                             activation, device)
        self.conv = nn.ModuleList(enc)
        make_heads(self._enc_out)
        self.dconv = nn.ModuleList(dec)
        n_levels = len(spiral_sizes) - 1
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("semantichuman_b200 models need a CUDA device; there is no CPU fallback")
        # per-level tables; level n_levels has spirals but no conv (models.py:71,90 loop to len-1)
        tables = [normalise_spiral(spirals[i]) for i in range(n_levels)]
        self._pD = [as_pool_matrix(D[i], dev) for i in range(n_levels)]
        self._pU = [as_pool_matrix(U[i], dev) for i in range(n_levels)]
        for i in range(n_levels):
            if self._pD[i].rows_in != tables[i].shape[0] or self._pU[i].rows_out != tables[i].shape[0]:
                raise ValueError(f"level {i}: D/U shapes do not match the spiral table")
        # Internal vertex order (invisible to callers): inside the trunks level l lives in the order perm[l] (new position i
        # holds the caller's vertex perm[l][i]; the dummy row stays last), chosen so that spiral neighbourhoods are
        # index-local.  Only tables change: spiral tables, D/U and the kept-row lists are relabelled, the level-0 tensor is
        # row-permuted on the way in and back on the way out, and the last level keeps the caller's order (FC layout).
        self._perm = self._level_orders(tables, n_levels) if reorder else None
        sel_rows = [pm.selection_cols if pm.is_selection else None for pm in self._pD]
        if self._perm is not None:
            full = [np.concatenate([q, [len(q)]]) for q in self._perm]           # + dummy
            pos = []
            for f in full:
                r = np.empty(len(f), np.int64)
                r[f] = np.arange(len(f))
                pos.append(r)
            tables = [pos[i][tables[i][full[i]]].astype(np.int32) for i in range(n_levels)]
            sel_rows = [None if sel_rows[i] is None else pos[i][sel_rows[i][full[i + 1]]] for i in range(n_levels)]
            self._pD = [self._pD[i].permuted(full[i + 1], full[i]) for i in range(n_levels)]
            self._pU = [self._pU[i].permuted(full[i], full[i + 1]) for i in range(n_levels)]
        self._perm_dev = None if self._perm is None else torch.as_tensor(full[0], dtype=torch.int32, device=dev)
        slab.inverse_perm(self._perm_dev)   # built now (and kept on the tensor): never inside a graph capture
        # a pool keeps the dummy row zero only if its last row is exactly e_dummy (main.py:190-191 builds it that way)
        d_ok = [pm.dummy_preserving for pm in self._pD]
        u_ok = [pm.dummy_preserving for pm in self._pU]
        geoms = [SpiralGeometry(tables[i], tables[i].shape[0], dev, dummy_row_grad=False) for i in range(n_levels)]
        # encoder plan: (conv index, geometry, pool-after or None)
        self._enc_plan = []
        for j, lvl in enumerate(self._enc_lvl):
            last_at_level = (j + 1 == len(self._enc_lvl)) or (self._enc_lvl[j + 1] != lvl)
            # conv 0 reads the caller's tensor (its dummy row is zero only by the dataset's convention); every later
            # encoder conv reads a masked conv output, possibly through a pool that maps dummy -> dummy
            first_at_level = (j == 0) or (self._enc_lvl[j - 1] != lvl)
            # input dummy row known zero: a masked conv output, directly or through a dummy-preserving pool / selection
            src_zero = (j > 0) and (not first_at_level or d_ok[lvl - 1])
            geoms[lvl] = geoms[lvl].with_flags(src_dummy_zero=src_zero, dummy_row_grad=not src_zero and j > 0)
            if not last_at_level:
                self._enc_plan.append((j, geoms[lvl], None))
            elif fuse_pool and sel_rows[lvl] is not None and sel_rows[lvl][-1] == geoms[lvl].rows_in - 1:
                # D is a row selection: evaluate the conv only at the kept vertices (+ dummy)
                self._enc_plan.append((j, geoms[lvl].restricted(sel_rows[lvl], dummy_row_grad=geoms[lvl].dummy_row_grad,
                                                                src_dummy_zero=src_zero), None))
            else:
                self._enc_plan.append((j, geoms[lvl], self._pD[lvl]))
        # decoder plan: (conv index, geometry, pool-before or None); only the very first decoder conv can see a live
        # dummy row (the FC output row carried by U[-1][-1,-1]=1) -> it alone computes that gradient
        self._dec_plan = []
        for j, lvl in enumerate(self._dec_lvl):
            first_at_level = (j == 0) or (self._dec_lvl[j - 1] != lvl)
            # decoder conv 0 sees the live FC row; all later decoder convs read masked outputs (through U: dummy -> dummy)
            src_zero = j > 0 and (not first_at_level or u_ok[lvl])
            g = geoms[lvl].with_flags(dummy_row_grad=not src_zero, src_dummy_zero=src_zero)
            self._dec_plan.append((j, g, self._pU[lvl] if first_at_level else None))
        self.compute_dtype = torch.float32

    def _level_orders(self, tables, n_levels):
        """perm[l] for l = 0..n_levels: reverse Cuthill-McKee at level 0; a coarser level inherits the order of the level
        above when D is a row selection (so that D/U stay local too), else gets its own; the last level is the identity."""
        try:
            perm = [locality_order(tables[0])]
        except ImportError:  # scipy missing: keep the caller's numbering (a performance matter only)
            return None
        for l in range(n_levels):
            v_next = self._pD[l].rows_out - 1
            if l + 1 == n_levels:
                perm.append(np.arange(v_next, dtype=np.int64))
            elif self._pD[l].is_selection and self._pD[l].selection_cols[-1] == self._pD[l].rows_in - 1:
                where = np.empty(len(perm[l]) + 1, np.int64)
                where[perm[l]] = np.arange(len(perm[l]))
                perm.append(np.argsort(where[self._pD[l].selection_cols[:-1]], kind="stable").astype(np.int64))
            else:
                perm.append(locality_order(tables[l + 1]))
        return perm

    def _planes(self):
        return 1 if self.compute_dtype == torch.bfloat16 else 2

    def _shadow_of(self, p):
        """bf16 copy of a master parameter, re-made only when the parameter changed (its autograd version) or moved.
        optim.Adam writes the shadows itself, in the same pass as the update."""
        if not hasattr(self, "_shadows"):
            self._shadows = {}
        ent = self._shadows.get(p)
        if ent is None or ent[0].device != p.device or ent[0].shape != p.shape:
            ent = [torch.empty_like(p, dtype=torch.bfloat16), -1]
            self._shadows[p] = ent
        if ent[1] != p._version:
            fn.cast_bf16(p.detach(), ent[0])
            ent[1] = p._version
        return ent[0]

    def shadow_map(self):
        """{parameter: [bf16 shadow, version]} of the parameters the bf16 mode reads through shadows (the FC layers), for
        optim.Adam(shadows=...)."""
        if not hasattr(self, "_shadows"):
            self._shadows = {}
        for name, p in self.named_parameters():
            if name.startswith("fc_latent_enc.") or name.startswith("fc_latent_dec."):
                self._shadow_of(p)
        return self._shadows

    def direct_grad_params(self):
        """Parameters whose gradient GEMM can write straight into a data-parallel bucket (dp.GradSync sinks): the two FC
        weights, in bf16 mode (fp32 mode runs them through nn.Linear's own autograd)."""
        if self.compute_dtype != torch.bfloat16:
            return []
        return [p for n, p in self.named_parameters() if n in ("fc_latent_enc.weight", "fc_latent_dec.weight")]

    def set_compute_dtype(self, dtype):
        """torch.float32 (default; every operand split into bf16 hi + lo, three tensor-core products per term, fp32
        accumulation: 1e-4 parity) or torch.bfloat16 (bf16 activations and operands, fp32 accumulation, fp32 master
        weights; 2e-2 parity)."""
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("compute dtype must be float32 or bfloat16")
        self.compute_dtype = dtype
        return self

    def _weight_images(self):
        """Operand images of every conv weight, current as of now (one launch if any weight changed since the last call)."""
        planes = self._planes()
        wi = getattr(self, "_wimg", None)
        weights = [c.conv.weight for c in list(self.conv) + list(self.dconv)]
        if wi is None or wi.planes != planes or len(wi.layers) != len(weights) or any(
                a is not b or a.device != wi.of(a)[0].device for (a, _), b in zip(wi.layers, weights)):
            wi = self._wimg = slab.WeightImages([(c.conv.weight, c.spiral_size) for c in list(self.conv) + list(self.dconv)],
                                                planes)
        wi.refresh()
        return wi

    def _encode_trunk(self, x):
        if not x.is_cuda:
            raise RuntimeError("semantichuman_b200 models need CUDA tensors; there is no CPU fallback")
        wi = self._weight_images()
        s = slab.from_rows(x, self._perm_dev, self._planes())  # caller's vertex order -> internal order, slab layout
        for j, geom, pm in self._enc_plan:
            c = self.conv[j]
            s = slab.spiral_conv(s, c.conv.weight, c.conv.bias, geom, c.activation_name, images=wi.of(c.conv.weight))
            if pm is not None:
                s = slab.pool(s, pm)
        return slab.to_rows(s, None, self.compute_dtype)  # last level: caller's order (FC layout of models.py:128)

    def _decode_slab(self, x):
        if not x.is_cuda:
            raise RuntimeError("semantichuman_b200 models need CUDA tensors; there is no CPU fallback")
        wi = self._weight_images()
        s = slab.from_rows(x, None, self._planes())
        for j, geom, pm in self._dec_plan:
            if pm is not None:
                s = slab.pool(s, pm)
            c = self.dconv[j]
            s = slab.spiral_conv(s, c.conv.weight, c.conv.bias, geom, c.activation_name, images=wi.of(c.conv.weight))
        return s

    def _decode_trunk(self, x):
        return slab.to_rows(self._decode_slab(x), self._perm_dev, torch.float32)  # internal order -> caller's order


class SpiralAutoencoder(_SpiralTrunk):
    """models.py:55-162."""

    def __init__(self, filters_enc, filters_dec, latent_size, sizes, spiral_sizes, spirals, D, U, device,
                 VAE_flag=False, activation='elu', fuse_pool=True, reorder=True):
        super().__init__()
        self.latent_size = latent_size
        self.VAE_flag = VAE_flag

        def heads(enc_out):
            self.fc_latent_enc = nn.Linear((sizes[-1] + 1) * enc_out, (2 if VAE_flag else 1) * latent_size)
            self.fc_latent_dec = nn.Linear(latent_size, (sizes[-1] + 1) * filters_dec[0][0])

        self._init_trunk(filters_enc, filters_dec, sizes, spiral_sizes, spirals, D, U, device, activation, fuse_pool,
                         heads, reorder)

    def _linear(self, layer, v):
        if self.compute_dtype == torch.float32:
            return layer(v)
        return fn.LinearShadowFn.apply(v, layer.weight, layer.bias, self._shadow_of(layer.weight), self._shadow_of(layer.bias))

    def encode(self, x, VAE_flag):
        bsize = x.size(0)
        x = self._encode_trunk(x)
        z = self._linear(self.fc_latent_enc, x.reshape(bsize, -1)).float()
        if VAE_flag:  # models.py:131-136
            self.z_mu = z[..., :self.latent_size]
            self.z_var = z[..., self.latent_size:]
            std = torch.exp(self.z_var / 2)
            eps = torch.randn_like(std)
            z = eps.mul(std).add_(self.z_mu)
        return z

    def decode(self, z):
        bsize = z.size(0)
        x = self._linear(self.fc_latent_dec, z.to(self.compute_dtype))
        return self._decode_trunk(x.view(bsize, self.sizes[-1] + 1, -1))

    def forward(self, x):
        z = self.encode(x, self.VAE_flag)
        return self.decode(z), z

    def reconstruction_loss(self, x, target=None):
        """F.l1_loss(self(x)[0], target) (train_funcs.py:501; target defaults to x) as ONE fused path for a training step: the
        latent code stays in the compute dtype between the two FC layers (bf16 -> fp32 -> bf16 is the identity, so the result
        equals forward()'s) and the loss is taken from the last SpiralConv's slab output (slab.l1_loss), so neither the
        row-major reconstruction nor its gradient is materialised.  Same loss and gradients as the unfused calls up to
        summation order."""
        if self.VAE_flag:
            return fn.l1_loss(self(x)[0], x if target is None else target)
        bsize = x.size(0)
        h = self._encode_trunk(x)
        z = self._linear(self.fc_latent_enc, h.reshape(bsize, -1))
        h = self._linear(self.fc_latent_dec, z)
        s = self._decode_slab(h.view(bsize, self.sizes[-1] + 1, -1))
        return slab.l1_loss(s, x if target is None else target, self._perm_dev)


class SpiralAutoencoder_multiz_partkps(_SpiralTrunk):
    """models.py:166-310 -- bone-guided variant: per-part shape codes + per-part keypoint (bone) codes."""

    def __init__(self, kps_index_list, vert_part_index_dict, filters_enc, filters_dec, latent_size, part_kps_latent_size,
                 sizes, spiral_sizes, spirals, D, U, device, VAE_flag=False, activation='elu', fuse_pool=True,
                 reorder=True, grouped_heads=True, skl_list=None):
        super().__init__()
        # models.py:169,285 read cfg.CONSTANTS.newskl_list at call time (27 bones by default, 31 under the shipped yaml
        # configs): pass `skl_list`, or assign `model.newskl_list` later -- kps_keep follows it
        self.newskl_list = DEFAULT_NEWSKL_LIST if skl_list is None else [list(b) for b in skl_list]
        self.kps_index_list = kps_index_list
        self.vert_part_index_dict = vert_part_index_dict
        self.part_kps_latent_size = part_kps_latent_size
        self.latent_size = latent_size
        self.VAE_flag = VAE_flag
        parts = [np.asarray(v) for v in vert_part_index_dict.values()]
        c_dec = filters_dec[0][0]
        out_lat = (2 if VAE_flag else 1) * latent_size

        def heads(c_enc):  # models.py:199-204
            self.fc_latent_enc_list = nn.ModuleList([nn.Linear(len(p) * c_enc, out_lat).to(device) for p in parts])
            self.fc_latent_dec_list = nn.ModuleList(
                [nn.Linear(latent_size + part_kps_latent_size, len(p) * c_dec).to(device) for p in parts])
            self.kps_enc_list = nn.ModuleList(
                [nn.Linear(len(k) * 3, part_kps_latent_size).to(device) for k in kps_index_list])

        self._init_trunk(filters_enc, filters_dec, sizes, spiral_sizes, spirals, D, U, device, activation, fuse_pool,
                         heads, reorder)
        dev = torch.device(device)
        self._part_idx = [torch.as_tensor(p, dtype=torch.long, device=dev) for p in parts]
        self._kps_idx = [torch.as_tensor(k, dtype=torch.long, device=dev) for k in kps_index_list]
        self._re_index = torch.as_tensor(np.concatenate(parts), dtype=torch.long, device=dev)
        # grouped-kernel form of the 17+17+17 per-part Linears: one launch per direction.  The decode form
        # needs the parts to tile the coarsest level exactly (they do: main.py:118-126 composes D to partition it).
        rows = sizes[-1] + 1
        self._g_enc = fn.GroupLayout(parts, rows, self._enc_out, out_lat, True, dev)
        self._g_dec = fn.GroupLayout(parts, rows, c_dec, latent_size + part_kps_latent_size, False, dev)
        n_kps = max(int(np.max(k)) for k in kps_index_list) + 1
        self._g_kps = fn.GroupLayout([np.asarray(k) for k in kps_index_list], n_kps, 3, part_kps_latent_size, True, dev)
        self._g_kps_by_rows = {n_kps: self._g_kps}
        covered = np.sort(np.concatenate(parts))
        self.grouped_heads = bool(grouped_heads and self._g_enc.supported() and self._g_dec.supported()
                                  and self._g_kps.supported() and self._g_enc.disjoint
                                  and np.array_equal(covered, np.arange(sizes[-1])))

    def kps_encode(self, kps):
        B = kps.shape[0]
        if self.grouped_heads and kps.is_cuda and not kps.requires_grad and kps.shape[1] >= self._g_kps.rows:
            lay = self._g_kps_by_rows.get(kps.shape[1])
            if lay is None:  # keypoint sets longer than the largest referenced index: same groups, wider row stride
                lay = fn.GroupLayout([k.cpu().numpy() for k in self._kps_idx], kps.shape[1], 3, self.part_kps_latent_size,
                                     True, kps.device)
                self._g_kps_by_rows[kps.shape[1]] = lay
            return fn.group_linear_gather(kps, self.kps_enc_list, lay)
        return torch.stack([self.kps_enc_list[k](kps[:, idx, :].reshape(B, -1)) for k, idx in enumerate(self._kps_idx)],
                           dim=1)

    def encode(self, x, kps, VAE_flag=None):
        bsize = x.size(0)
        x = self._encode_trunk(x).float()
        if self.grouped_heads:
            z = fn.group_linear_gather(x, self.fc_latent_enc_list, self._g_enc)
        else:
            z = torch.stack([self.fc_latent_enc_list[k](x[:, idx, :].reshape(bsize, -1))
                             for k, idx in enumerate(self._part_idx)], dim=1)
        return z, self.kps_encode(kps), x[:, -1:, :]

    def decode(self, z, z_part_kps, dummy):
        bsize = z.size(0)
        zz = torch.cat([z, z_part_kps], dim=2)
        if self.grouped_heads:
            return self._decode_trunk(fn.group_linear_scatter(zz, self.fc_latent_dec_list, dummy, self._g_dec))
        pieces = [self.fc_latent_dec_list[k](zz[:, k, :]) for k in range(z.shape[1])]
        x = torch.cat(pieces, dim=1).view(bsize, self.sizes[-1], -1)
        # models.py:270-272: x[:, re_index] = x[:, arange]  (rows arrive in part-concatenated order)
        x = torch.zeros_like(x).index_copy(1, self._re_index, x)
        x = torch.cat([x, dummy.to(x.dtype).expand(bsize, -1, -1)], dim=1)
        return self._decode_trunk(x)

    @property
    def kps_keep(self):
        return [i for i in range(len(self.newskl_list) + 4) if i not in (3, 13, 14)]

    def kps2skl(self, kps_tmp):
        """models.py:284-304: keypoints -> (unit direction, length) per bone of ``newskl_list``."""
        skl_list = self.newskl_list
        if kps_tmp.shape[1] == len(skl_list) + 4:
            kps = copy.deepcopy(kps_tmp)
        else:
            keep = [i for i in range(len(skl_list) + 4) if i not in (3, 13, 14)]
            kps = torch.zeros((kps_tmp.shape[0], len(skl_list) + 4, 3), device=kps_tmp.device)
            kps[:, keep, :] = kps_tmp
        skl = torch.zeros((kps.shape[0], len(skl_list), 4), device=kps.device)
        for i, bone in enumerate(skl_list):
            tail = kps[:, bone[1], :] if len(bone) == 2 else (kps[:, bone[1], :] + kps[:, bone[2], :]) / 2
            vec = kps[:, bone[0], :] - tail
            length = torch.sqrt(torch.sum(vec ** 2, dim=1))
            skl[:, i, :3] = vec / length[:, None]
            skl[:, i, -1] = length
        return skl

    def forward(self, x, kps):
        z, z_part_kps, dummy = self.encode(x, kps, self.VAE_flag)
        return self.decode(z, z_part_kps, dummy), z, z_part_kps
