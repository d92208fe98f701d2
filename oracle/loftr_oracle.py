"""ORACLE (test infrastructure only — never imported by the product path).

Plain PyTorch fp32 CPU restatement of the LoFTR inference path that the reference's
``LoFTR_matcher`` runs (etc/feature_matching_baselines/matchers.py:12-59 ->
etc/feature_matching_baselines/LoFTR/src/loftr/loftr.py:29-75), written functionally over a
reference-format ``state_dict`` so the same weights feed the reference module, this oracle and the
CUDA engine. Dual-softmax matcher, ``temp_bug_fix=False`` position encoding
(utils/cvpr_ds_config.py:28,34).

Pinned against the reference itself (imported from /root/reference in the build container) by
``tests/test_oracle_vs_reference.py`` and the committed vectors in ``tests/golden/``.
"""
import math

import torch
import torch.nn.functional as F

CONFIG = {
    "thr": 0.2, "border_rm": 2, "temperature": 0.1,          # cvpr_ds_config.py:31-35
    "coarse_layers": ["self", "cross"] * 4, "coarse_heads": 8,  # :22-27
    "fine_layers": ["self", "cross"], "fine_heads": 8,           # :44-48
    "window": 5,                                                  # :12
}


# ------------------------------------------------------------------------------ backbone
def _bn(x, sd, pre):
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"],
                        sd[pre + ".weight"], sd[pre + ".bias"], False, 0.0, 1e-5)


def _basic_block(x, sd, pre, stride):
    """backbone/resnet_fpn.py:15-40."""
    y = F.relu(_bn(F.conv2d(x, sd[pre + ".conv1.weight"], None, stride, 1), sd, pre + ".bn1"))
    y = _bn(F.conv2d(y, sd[pre + ".conv2.weight"], None, 1, 1), sd, pre + ".bn2")
    if stride != 1:
        x = _bn(F.conv2d(x, sd[pre + ".downsample.0.weight"], None, stride, 0), sd,
                pre + ".downsample.1")
    return F.relu(x + y)


def backbone(img, sd, pre="backbone"):
    """ResNetFPN_8_2.forward, backbone/resnet_fpn.py:100-119. img [N,1,H,W] -> (1/8, 1/2) maps."""
    x0 = F.relu(_bn(F.conv2d(img, sd[pre + ".conv1.weight"], None, 2, 3), sd, pre + ".bn1"))
    x1 = _basic_block(_basic_block(x0, sd, pre + ".layer1.0", 1), sd, pre + ".layer1.1", 1)
    x2 = _basic_block(_basic_block(x1, sd, pre + ".layer2.0", 2), sd, pre + ".layer2.1", 1)
    x3 = _basic_block(_basic_block(x2, sd, pre + ".layer3.0", 2), sd, pre + ".layer3.1", 1)
    x3_out = F.conv2d(x3, sd[pre + ".layer3_outconv.weight"])
    up = F.interpolate(x3_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x2_out = F.conv2d(x2, sd[pre + ".layer2_outconv.weight"]) + up
    x2_out = F.conv2d(x2_out, sd[pre + ".layer2_outconv2.0.weight"], None, 1, 1)
    x2_out = F.leaky_relu(_bn(x2_out, sd, pre + ".layer2_outconv2.1"), 0.01)
    x2_out = F.conv2d(x2_out, sd[pre + ".layer2_outconv2.3.weight"], None, 1, 1)
    up = F.interpolate(x2_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x1_out = F.conv2d(x1, sd[pre + ".layer1_outconv.weight"]) + up
    x1_out = F.conv2d(x1_out, sd[pre + ".layer1_outconv2.0.weight"], None, 1, 1)
    x1_out = F.leaky_relu(_bn(x1_out, sd, pre + ".layer1_outconv2.1"), 0.01)
    x1_out = F.conv2d(x1_out, sd[pre + ".layer1_outconv2.3.weight"], None, 1, 1)
    return x3_out, x1_out


# ------------------------------------------------------------------------------ position encoding
def position_encoding(d_model, h, w):
    """utils/position_encoding.py:22-38 with temp_bug_fix=False: the divisor expression
    ``-math.log(10000.0) / d_model // 2`` floor-divides to -1.0, so div_term = exp(-2k)."""
    pe = torch.zeros((d_model, h, w))
    y_pos = torch.ones((h, w)).cumsum(0).float().unsqueeze(0)
    x_pos = torch.ones((h, w)).cumsum(1).float().unsqueeze(0)
    div = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / d_model // 2))
    div = div[:, None, None]
    pe[0::4] = torch.sin(x_pos * div)
    pe[1::4] = torch.cos(x_pos * div)
    pe[2::4] = torch.sin(y_pos * div)
    pe[3::4] = torch.cos(y_pos * div)
    return pe


# ------------------------------------------------------------------------------ transformer
def linear_attention(q, k, v, eps=1e-6):
    """loftr_module/linear_attention.py:20-47. q [N,L,H,D], k/v [N,S,H,D]."""
    Q = F.elu(q) + 1
    K = F.elu(k) + 1
    S = v.size(1)
    v = v / S
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + eps)
    return torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * S


def encoder_layer(x, src, sd, pre, nhead):
    """LoFTREncoderLayer.forward, loftr_module/transformer.py:35-58."""
    n, _, c = x.shape
    d = c // nhead
    q = F.linear(x, sd[pre + ".q_proj.weight"]).view(n, -1, nhead, d)
    k = F.linear(src, sd[pre + ".k_proj.weight"]).view(n, -1, nhead, d)
    v = F.linear(src, sd[pre + ".v_proj.weight"]).view(n, -1, nhead, d)
    msg = F.linear(linear_attention(q, k, v).reshape(n, -1, c), sd[pre + ".merge.weight"])
    msg = F.layer_norm(msg, (c,), sd[pre + ".norm1.weight"], sd[pre + ".norm1.bias"])
    msg = F.linear(torch.cat([x, msg], dim=2), sd[pre + ".mlp.0.weight"])
    msg = F.linear(F.relu(msg), sd[pre + ".mlp.2.weight"])
    msg = F.layer_norm(msg, (c,), sd[pre + ".norm2.weight"], sd[pre + ".norm2.bias"])
    return x + msg


def transformer(f0, f1, sd, pre, names, nhead):
    """LocalFeatureTransformer.forward, transformer.py:80-101: in 'cross' layers feat1 attends to
    the already-updated feat0."""
    for i, name in enumerate(names):
        p = f"{pre}.layers.{i}"
        if name == "self":
            f0 = encoder_layer(f0, f0, sd, p, nhead)
            f1 = encoder_layer(f1, f1, sd, p, nhead)
        else:
            f0 = encoder_layer(f0, f1, sd, p, nhead)
            f1 = encoder_layer(f1, f0, sd, p, nhead)
    return f0, f1


# ------------------------------------------------------------------------------ coarse matching
def coarse_conf(f0, f1, temperature):
    """utils/coarse_matching.py:109-119 (dual-softmax)."""
    c = f0.shape[-1]
    f0 = f0 / c ** 0.5
    f1 = f1 / c ** 0.5
    sim = torch.einsum("nlc,nsc->nls", f0, f1) / temperature
    return F.softmax(sim, 1) * F.softmax(sim, 2), sim


def coarse_select(conf, h0, w0, h1, w1, thr, border):
    """CoarseMatching.get_coarse_match, coarse_matching.py:175-196 (inference branch).
    Returns b_ids, i_ids, j_ids (int64) and mconf."""
    n = conf.shape[0]
    mask = (conf > thr).view(n, h0, w0, h1, w1).clone()
    if border > 0:
        b = border
        mask[:, :b] = False
        mask[:, :, :b] = False
        mask[:, :, :, :b] = False
        mask[:, :, :, :, :b] = False
        mask[:, -b:] = False
        mask[:, :, -b:] = False
        mask[:, :, :, -b:] = False
        mask[:, :, :, :, -b:] = False
    mask = mask.view(n, h0 * w0, h1 * w1)
    mask = mask * (conf == conf.max(dim=2, keepdim=True)[0]) * (conf == conf.max(dim=1, keepdim=True)[0])
    mask_v, all_j = mask.max(dim=2)
    b_ids, i_ids = torch.where(mask_v)
    j_ids = all_j[b_ids, i_ids]
    return b_ids, i_ids, j_ids, conf[b_ids, i_ids, j_ids]


# ------------------------------------------------------------------------------ fine level
def fine_preprocess(ff0, ff1, fc0, fc1, b_ids, i_ids, j_ids, sd, stride, W=5):
    """FinePreprocess.forward, loftr_module/fine_preprocess.py:29-59."""
    cf = ff0.shape[1]
    if b_ids.numel() == 0:
        e = torch.empty(0, W * W, cf, device=ff0.device)
        return e, e
    def unfold(f):
        u = F.unfold(f, kernel_size=(W, W), stride=stride, padding=W // 2)      # [n, c*ww, l]
        n, _, l = u.shape
        return u.view(n, cf, W * W, l).permute(0, 3, 2, 1)                      # n l ww c
    w0 = unfold(ff0)[b_ids, i_ids]
    w1 = unfold(ff1)[b_ids, j_ids]
    cw = torch.cat([fc0[b_ids, i_ids], fc1[b_ids, j_ids]], 0)
    cw = F.linear(cw, sd["fine_preprocess.down_proj.weight"], sd["fine_preprocess.down_proj.bias"])
    cat = torch.cat([torch.cat([w0, w1], 0), cw[:, None, :].expand(-1, W * W, -1)], -1)
    out = F.linear(cat, sd["fine_preprocess.merge_feat.weight"], sd["fine_preprocess.merge_feat.bias"])
    return torch.chunk(out, 2, dim=0)


def fine_matching(f0, f1, mkpts0_c, mkpts1_c, scale, W=5):
    """FineMatching.forward / get_fine_match, utils/fine_matching.py:15-74 (kornia's
    spatial_expectation2d over a normalised [-1,1] grid restated inline)."""
    M, WW, C = f0.shape
    if M == 0:
        return mkpts0_c, mkpts1_c, torch.empty(0, 3, device=f0.device)
    sim = torch.einsum("mc,mrc->mr", f0[:, WW // 2, :], f1)
    heat = torch.softmax(sim / C ** 0.5, dim=1)                                # [M, WW]
    lin = torch.linspace(-1, 1, W, device=f0.device)
    gx = lin[None, :].expand(W, W).reshape(-1)                                 # x varies fastest
    gy = lin[:, None].expand(W, W).reshape(-1)
    ex = (heat * gx).sum(1)
    ey = (heat * gy).sum(1)
    coords = torch.stack([ex, ey], 1)
    grid = torch.stack([gx, gy], 1)
    var = torch.sum(grid[None] ** 2 * heat[:, :, None], dim=1) - coords ** 2
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    mk1 = mkpts1_c + coords * (W // 2) * scale
    return mkpts0_c, mk1, torch.cat([coords, std[:, None]], -1)


# ------------------------------------------------------------------------------ end to end
def loftr_forward(img0, img1, sd, cfg=None, return_intermediates=False):
    """LoFTR.forward, loftr.py:29-75 for equal-sized inputs [N,1,H,W] (multiples of 8).
    Returns dict with mkpts0_f, mkpts1_f, mconf, b_ids, i_ids, j_ids (+ intermediates)."""
    cfg = {**CONFIG, **(cfg or {})}
    n = img0.shape[0]
    fc, ff = backbone(torch.cat([img0, img1], 0), sd)
    h_c, w_c = fc.shape[2:]
    h_f = ff.shape[2]
    pe = position_encoding(fc.shape[1], h_c, w_c).to(fc.device)     # device-agnostic: bench.py also times it on the GPU
    fc = (fc + pe[None]).flatten(2).transpose(1, 2)                            # n (h w) c
    fc0, fc1 = fc[:n], fc[n:]
    ff0, ff1 = ff[:n], ff[n:]
    fc0, fc1 = transformer(fc0, fc1, sd, "loftr_coarse", cfg["coarse_layers"], cfg["coarse_heads"])
    conf, sim = coarse_conf(fc0, fc1, cfg["temperature"])
    b_ids, i_ids, j_ids, mconf = coarse_select(conf, h_c, w_c, h_c, w_c, cfg["thr"], cfg["border_rm"])
    scale_c = img0.shape[2] / h_c
    mk0c = torch.stack([i_ids % w_c, i_ids // w_c], 1) * scale_c
    mk1c = torch.stack([j_ids % w_c, j_ids // w_c], 1) * scale_c
    w0, w1 = fine_preprocess(ff0, ff1, fc0, fc1, b_ids, i_ids, j_ids, sd, h_f // h_c, cfg["window"])
    if w0.shape[0] != 0:
        w0, w1 = transformer(w0, w1, sd, "loftr_fine", cfg["fine_layers"], cfg["fine_heads"])
    mk0f, mk1f, expec = fine_matching(w0, w1, mk0c, mk1c, img0.shape[2] / h_f, cfg["window"])
    out = {"mkpts0_f": mk0f, "mkpts1_f": mk1f, "mconf": mconf, "b_ids": b_ids, "i_ids": i_ids,
           "j_ids": j_ids}
    if return_intermediates:
        out.update({"feat_c0": fc0, "feat_c1": fc1, "feat_f0": ff0, "feat_f1": ff1, "conf": conf,
                    "sim": sim, "win0": w0, "win1": w1, "expec_f": expec})
    return out


# ------------------------------------------------------------------------------ synthetic weights
def make_state_dict(seed=0):
    """Deterministic synthetic LoFTR weights in the reference's state_dict format (shared with the
    product package so that the oracle, the reference module and the CUDA engine see the same bits)."""
    from mfr_b200.loftr import synthetic_state_dict
    return synthetic_state_dict(seed)
