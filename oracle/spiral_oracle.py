"""CPU oracle for the SemanticHuman spiral-autoencoder hot path.   *** TEST INFRASTRUCTURE ***

This file restates, as plain functions over CPU torch tensors (any float dtype; float64 is the tie-breaker),
the algorithm of the reference's ``models.py`` and of the loss lines of ``train_funcs.py``.  It exists so that
the CUDA path can be checked on a box that does not have ``/root/reference``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it;
the product package ``semantichuman_b200`` never does (tests/test_capi_cpu.py::test_product_never_imports_oracle enforces that).

Pinning: the reference has no tests or golden vectors of its own, so the oracle is pinned
against outputs of the reference itself, generated in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference/models.py`` unmodified) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them.

The hot-path arithmetic of the reference lives in ATen (torch 1.10 pinned, README.md:23; torch 2.11 here):
``index`` / ``addmm`` / ``elu`` / ``mul`` / ``bmm`` / ``l1_loss``.  The same ATen ops are used below, through a
functional formulation; gradients come from autograd exactly as in the reference's ``loss.backward()``.
"""
import numpy as np
import torch
import torch.nn.functional as F

ACTIVATIONS = {
    # models.py:19-32
    "relu": torch.relu,
    "elu": F.elu,
    "leaky_relu": lambda v: F.leaky_relu(v, 0.02),
    "sigmoid": torch.sigmoid,
    "tanh": torch.tanh,
    "identity": lambda v: v,
}


def spiral_conv(x, spiral_idx, weight, bias, act="elu"):
    """models.py:34-53.  x (B, V+1, Cin); spiral_idx (V+1, S) or (1|B, V+1, S) integer with -1 == dummy row;
    weight (Cout, S*Cin) with k = s*Cin + c; returns (B, V+1, Cout) with the last row zeroed."""
    if act not in ACTIVATIONS:
        raise NotImplementedError(act)  # models.py:31-32
    idx = spiral_idx if spiral_idx.dim() == 2 else spiral_idx[0]
    B, V1, Cin = x.shape
    S = idx.shape[1]
    rows = idx.long() % V1  # Python negative indexing: -1 -> V1-1  (models.py:42)
    gathered = x[:, rows.reshape(-1), :].reshape(B * V1, S * Cin)  # models.py:42
    out = F.linear(gathered, weight, bias)  # models.py:45
    out = ACTIVATIONS[act](out).reshape(B, V1, -1)  # models.py:46,48
    mask = torch.ones(1, V1, 1, dtype=x.dtype)
    mask[0, -1, 0] = 0.0  # models.py:49-50
    return out * mask  # models.py:51


def pool(P, x):
    """models.py:127 / :148 -- dense (1, Vout+1, Vin+1) sampling matrix times (B, Vin+1, C)."""
    return torch.matmul(P.to(x.dtype), x)


def conv_plan(filters_enc, filters_dec, n_levels):
    """Layer list of SpiralAutoencoder.__init__ (models.py:69-113) as tuples
    (stack, level, in_c, out_c, last_is_identity)."""
    enc, dec = [], []
    c = filters_enc[0][0]
    for i in range(n_levels):
        if filters_enc[1][i]:
            enc.append((i, c, filters_enc[1][i], False))
            c = filters_enc[1][i]
        enc.append((i, c, filters_enc[0][i + 1], False))
        c = filters_enc[0][i + 1]
    c = filters_dec[0][0]
    for i in range(n_levels):
        lvl = n_levels - 1 - i  # spiral_sizes[-2-i]
        if i != n_levels - 1:
            dec.append((lvl, c, filters_dec[0][i + 1], False))
            c = filters_dec[0][i + 1]
            if filters_dec[1][i + 1]:
                dec.append((lvl, c, filters_dec[1][i + 1], False))
                c = filters_dec[1][i + 1]
        else:
            if filters_dec[1][i + 1]:
                dec.append((lvl, c, filters_dec[0][i + 1], False))
                c = filters_dec[0][i + 1]
                dec.append((lvl, c, filters_dec[1][i + 1], True))
                c = filters_dec[1][i + 1]
            else:
                dec.append((lvl, c, filters_dec[0][i + 1], True))
                c = filters_dec[0][i + 1]
    return enc, dec


def _run_stack(params, prefix, plan, x, spirals, act, level_hook):
    """Apply the convs of `plan` in order; `level_hook(level_index_in_stack, x)` runs between levels."""
    j = 0
    levels = []
    for item in plan:
        if item[0] not in levels:
            levels.append(item[0])
    for lvl in levels:
        x = level_hook("pre", lvl, x)
        for (l, _cin, _cout, ident) in plan:
            if l != lvl:
                continue
            x = spiral_conv(x, spirals[lvl], params[f"{prefix}.{j}.conv.weight"], params[f"{prefix}.{j}.conv.bias"],
                            "identity" if ident else act)
            j += 1
        x = level_hook("post", lvl, x)
    return x


def ae_encode_trunk(params, x, filters_enc, filters_dec, spirals, D, act="elu"):
    """Encoder conv + down-pool stack shared by both models (models.py:121-127, 244-250)."""
    n_levels = len(spirals) - 1
    enc, _ = conv_plan(filters_enc, filters_dec, n_levels)
    return _run_stack(params, "conv", enc, x, spirals, act,
                      lambda when, lvl, v: pool(D[lvl], v) if when == "post" else v)


def ae_decode_trunk(params, x, filters_enc, filters_dec, spirals, U, act="elu"):
    """Decoder up-pool + conv stack shared by both models (models.py:147-153, 275-281)."""
    n_levels = len(spirals) - 1
    _, dec = conv_plan(filters_enc, filters_dec, n_levels)
    return _run_stack(params, "dconv", dec, x, spirals, act,
                      lambda when, lvl, v: pool(U[lvl], v) if when == "pre" else v)


def autoencoder_forward(params, x, filters_enc, filters_dec, sizes, spirals, D, U, act="elu"):
    """SpiralAutoencoder.forward (models.py:115-162), VAE_flag False.  Returns (x_hat, z)."""
    B = x.shape[0]
    h = ae_encode_trunk(params, x, filters_enc, filters_dec, spirals, D, act)
    z = F.linear(h.reshape(B, -1), params["fc_latent_enc.weight"], params["fc_latent_enc.bias"])  # :129-130
    h = F.linear(z, params["fc_latent_dec.weight"], params["fc_latent_dec.bias"])  # :144
    h = h.reshape(B, sizes[-1] + 1, -1)  # :145
    return ae_decode_trunk(params, h, filters_enc, filters_dec, spirals, U, act), z


def multiz_kps_encode(params, kps, kps_index_list):
    """models.py:233-236."""
    B = kps.shape[0]
    return torch.stack([F.linear(kps[:, idx, :].reshape(B, -1), params[f"kps_enc_list.{k}.weight"],
                                 params[f"kps_enc_list.{k}.bias"]) for k, idx in enumerate(kps_index_list)], dim=1)


def multiz_encode(params, x, kps, kps_index_list, part_index_lists, filters_enc, filters_dec, spirals, D, act="elu"):
    """models.py:238-263.  Returns (z (B,P,L), z_kps (B,P,Lk), dummy (B,1,C))."""
    B = x.shape[0]
    h = ae_encode_trunk(params, x, filters_enc, filters_dec, spirals, D, act)
    z = torch.stack([F.linear(h[:, torch.as_tensor(p), :].reshape(B, -1), params[f"fc_latent_enc_list.{k}.weight"],
                              params[f"fc_latent_enc_list.{k}.bias"]) for k, p in enumerate(part_index_lists)], dim=1)
    return z, multiz_kps_encode(params, kps, kps_index_list), h[:, -1:, :]


def multiz_decode(params, z, z_kps, dummy, part_index_lists, filters_enc, filters_dec, sizes, spirals, U, act="elu"):
    """models.py:265-282."""
    B = z.shape[0]
    pieces = [F.linear(torch.cat([z[:, k, :], z_kps[:, k, :]], dim=1), params[f"fc_latent_dec_list.{k}.weight"],
                       params[f"fc_latent_dec_list.{k}.bias"]) for k in range(z.shape[1])]
    h = torch.cat(pieces, dim=1).reshape(B, sizes[-1], -1)  # :269, rows in part-concatenated order
    re_index = torch.as_tensor(np.concatenate([np.asarray(p) for p in part_index_lists]))
    placed = torch.zeros_like(h).index_copy(1, re_index, h)  # :270-272  x[:, re_index] = x[:, arange]
    h = torch.cat([placed, dummy], dim=1)  # :273
    return ae_decode_trunk(params, h, filters_enc, filters_dec, spirals, U, act)


def multiz_forward(params, x, kps, kps_index_list, part_index_lists, filters_enc, filters_dec, sizes, spirals, D, U,
                   act="elu"):
    """models.py:306-310.  Returns (x_hat, z, z_kps)."""
    z, zk, dummy = multiz_encode(params, x, kps, kps_index_list, part_index_lists, filters_enc, filters_dec, spirals, D,
                                 act)
    return multiz_decode(params, z, zk, dummy, part_index_lists, filters_enc, filters_dec, sizes, spirals, U, act), z, zk


def l1_loss(a, b):
    """F.l1_loss as used at train_funcs.py:135,501 (mean over every element incl. the dummy row)."""
    return (a - b).abs().mean()


def zpart_reg(z, measure, P, Q, relative=True):
    """train_funcs.py:145-152."""
    m = torch.sqrt(torch.sum(z ** 2, dim=2))
    if relative:
        return (m[:, P] / measure[:, Q] - 1.0).abs().mean()
    return (m[:, P] - measure[:, Q]).abs().mean()


# ---------------------------------------------------------------------------------------------- index oracles
def angle_weights(v, kps, part_index_lists, skl_list):
    """utils_SH.py:442-478 (angle_skl): per part, the angle in degrees (0..90) between every vertex pair direction
    v_i - v_j and the part's bone direction; |cos| clamped to [0, 1], NaN (i == j) -> cos = 1 -> 0 degrees.
    v (B, V, 3), kps (B, n_kps, 3); returns a list of (B, n_p, n_p) tensors."""
    out = []
    for idx, bone in zip(part_index_lists, skl_list):
        vp = v[:, idx, :]
        d = vp[:, :, None, :] - vp[:, None, :, :]                                   # :449
        if len(bone) == 2:                                                          # :450-453
            kd = kps[:, bone[0], :] - kps[:, bone[1], :]
        else:
            kd = kps[:, bone[0], :] - (kps[:, bone[1], :] + kps[:, bone[2], :]) / 2
        kd = kd[:, None, None, :]
        dm = torch.sqrt(torch.sum(d * d, dim=-1))                                   # :454
        km = torch.sqrt(torch.sum(kd * kd, dim=-1))                                 # :458
        cos = torch.abs(torch.sum(d * kd, dim=-1) / (dm * km))                      # :459-460
        cos = torch.where(torch.isnan(cos), torch.ones_like(cos), cos).float()      # :462
        cos = cos.clamp(0.0, 1.0)                                                   # :463-464
        out.append(torch.arccos(cos) * 180 / np.pi)                                 # :468
    return out


def euclidean_dist_matrix(x):
    """utils_distance.py:366-376 (calc_euclidean_dist_matrix): sqrt(relu(|x_i|^2 - 2 x_i.x_j + |x_j|^2))."""
    r = torch.sum(x ** 2, dim=2).unsqueeze(2)
    return F.relu(r - 2 * torch.bmm(x, x.transpose(2, 1)) + r.transpose(2, 1)) ** 0.5


def pair_distance_loss(tx, rec, kps, part_index_lists, skl_list, w_mode="linear", w_threshold=0.8, leaf_parts=(),
                       relative=True, part_weights=None, scale=None):
    """The orientation-adaptive pairwise-distance loss of train_funcs.py:243-284 (== :353-389 without `scale`).

    Per part p: w = f(angle)/... by `w_mode` (all-one for parts in `leaf_parts`, :259-267), diagonal zeroed (:268-269);
    De / De_r = pairwise distances of the ground truth / the reconstruction (:246-247), De optionally scaled per sample
    (`a[:, ...]`, :248-249); over the entries with w*De != 0 (:272): mean |w*De_r/De - w| (relat_flag, :276) or
    mean |w*De_r - w*De| (:274); summed with `part_weights` (default 1/K, :252-253)."""
    K = len(part_index_lists)
    ang = angle_weights(tx, kps, part_index_lists, skl_list)
    total = 0.0
    for p, idx in enumerate(part_index_lists):
        De = euclidean_dist_matrix(tx[:, idx, :])
        De_r = euclidean_dist_matrix(rec[:, idx, :])
        if scale is not None:
            De = De * scale[:, p][:, None, None]
        if w_mode == "all_one" or p in leaf_parts:
            w = torch.ones_like(ang[p])
        elif w_mode == "linear":
            w = ang[p].float() / 90
        elif w_mode == "sin":
            w = torch.sin(ang[p].float() / 180 * torch.pi)
        elif w_mode == "threshold":
            w = ang[p].float() / 90
            w = torch.where(w < w_threshold, torch.zeros_like(w), w)
        else:
            raise NotImplementedError(w_mode)
        w = w - torch.diag_embed(torch.diagonal(w, dim1=1, dim2=2))
        nz = torch.where((w * De) != 0)
        if relative:
            lp = F.l1_loss(w[nz] * De_r[nz].float() / De[nz], w[nz] * torch.ones_like(w[nz]))
        else:
            lp = F.l1_loss(w[nz] * De_r[nz].float(), w[nz] * De[nz])
        total = total + (1.0 / K if part_weights is None else part_weights[p]) * lp
    return total


def bone_guided_step_loss(params, tx, tx_interp, tx_exc, measure, J, kps_keep, kps_index_list, part_index_lists, skl_list, P, Q,
                          factor, filters_enc, filters_dec, sizes, spirals, D, U, weights, w_mode="linear", w_threshold=0.8,
                          leaf_parts=(0, 7, 10, 13, 16), relative=True, act="elu", part_index_lists_fine=None):
    """The loss of ONE bone-guided training step, train_funcs.py:128-392, for edit_mode 'equal' (:213-223), exc_mode 'ori_m'
    (:296-300), w_part_mode '1/K', without the per-sample edge / volume regularisers (:136-143, :322-332, separate oracle
    functions in aux losses).  `factor` is the value the loop draws at :222 (passed in so that the oracle is deterministic).
    `part_index_lists` are the part vertex lists at the COARSEST level (the model's heads, models.py:199-204);
    `part_index_lists_fine` those at level 0 (vert_part_index_dict of the loop, used by the distance loss; default: the same).
    Returns (total, dict of the six terms)."""
    fine = part_index_lists if part_index_lists_fine is None else part_index_lists_fine
    fw = dict(filters_enc=filters_enc, filters_dec=filters_dec, spirals=spirals, act=act)
    keep = torch.as_tensor(kps_keep)
    # :130-135 reconstruction pass
    kps_gt = torch.matmul(J, tx[:, :-1, :])
    tx_hat, z, _ = multiz_forward(params, tx, kps_gt[:, keep], kps_index_list, part_index_lists, filters_enc, filters_dec, sizes,
                                  spirals, D, U, act)
    terms = {"rec": l1_loss(tx, tx_hat), "zpartreg": zpart_reg(z, measure, P, Q, relative)}                  # :135, :144-152
    # :160-228 interpolation pass: scale the non-leaf part codes by one factor
    kps_i = torch.matmul(J, tx_interp[:, :-1, :])
    new_kps = kps_i[:, keep]                                                                                  # :218-219
    lat, lat_k, dummy = multiz_encode(params, tx_interp, new_kps, kps_index_list, part_index_lists, D=D, **fw)   # :225
    a = torch.ones(tx_interp.shape[0], len(P), dtype=tx.dtype) * factor                                       # :223
    lat = lat.clone()
    for k, v in enumerate(P):
        lat[:, v, :] = lat[:, v, :] * a[:, k][:, None]                                                        # :226-227
    rec_i = multiz_decode(params, lat, lat_k, dummy, part_index_lists, filters_enc, filters_dec, sizes, spirals, U, act)
    terms["interp_kps"] = l1_loss(torch.matmul(J, rec_i[:, :-1, :])[:, keep], new_kps)                        # :230-233
    scale = torch.ones(tx_interp.shape[0], len(part_index_lists), dtype=tx.dtype)
    scale[:, list(P)] = a
    terms["interp_euc"] = pair_distance_loss(tx_interp[:, :-1, :], rec_i[:, :-1, :], kps_i, fine, skl_list, w_mode,
                                             w_threshold, leaf_parts, relative, scale=scale)                  # :235-284
    # :286-389 exchange pass: every sample is encoded with the keypoints of its mirror sample in the batch
    kps_e = torch.matmul(J, tx_exc[:, :-1, :])
    new_kps_e = torch.flip(kps_e, dims=[0])[:, keep]                                                          # :298-300
    lat, lat_k, dummy = multiz_encode(params, tx_exc, new_kps_e, kps_index_list, part_index_lists, D=D, **fw)  # :319
    rec_e = multiz_decode(params, lat, lat_k, dummy, part_index_lists, filters_enc, filters_dec, sizes, spirals, U, act)
    terms["exc_kps"] = l1_loss(torch.matmul(J, rec_e[:, :-1, :])[:, keep], new_kps_e)                         # :334-341
    terms["exc_euc"] = pair_distance_loss(tx_exc[:, :-1, :], rec_e[:, :-1, :], kps_e, fine, skl_list, w_mode,
                                          w_threshold, leaf_parts, relative)                                  # :343-389
    total = sum(weights[k] * v for k, v in terms.items())
    return total, terms


def normalise_spiral(spiral_idx, rows_in=None):
    """-1 -> rows_in-1 (models.py:42's negative index), int32, 2-D."""
    a = np.asarray(spiral_idx)
    if a.ndim == 3:
        a = a[0]
    a = a.astype(np.int64)
    n = a.shape[0] if rows_in is None else rows_in
    return np.where(a < 0, a + n, a).astype(np.int32)


def inverse_spiral_csr(table, rows_in):
    """stable argsort of the flattened normalised table."""
    flat = np.asarray(table).reshape(-1).astype(np.int64)
    slots = np.argsort(flat, kind="stable").astype(np.int32)
    rowptr = np.zeros(rows_in + 1, np.int32)
    np.add.at(rowptr, flat + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), slots


def dense_to_csr(dense):
    d = np.asarray(dense)
    r, c = np.nonzero(d)
    rowptr = np.zeros(d.shape[0] + 1, np.int32)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), c.astype(np.int32), d[r, c].astype(np.float32)
