"""CPU oracle for the RECOMBINER hot path -- TEST INFRASTRUCTURE ONLY.

This module is a functional, explicit-noise restatement (torch CPU, f32 model /
f64 REC) of the reference algorithm for the path named in BASELINE.json
`north_star`.  It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product (`recombiner_b200/`) never imports anything from `oracle/`.

Parity pinning: the reference ships no tests or golden vectors,
so this restatement is pinned against outputs of the *unmodified reference run
in the build container* -- `oracle/make_golden.py` imports `/root/reference`
and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every
function below against those fixtures.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).  Noise is always an explicit argument (the reference draws
it from the global torch RNG, test_model.py:284-285,303; utils.py:148,181,189,196).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

LN2 = math.log(2.0)


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #
def std_transform(raw: torch.Tensor) -> torch.Tensor:
    """sigma = softplus(raw)/6, threshold 20 (test_model.py:101, prior_model.py:88)."""
    return F.softplus(raw, beta=1, threshold=20) / 6


def std_transform_inverse(sigma: torch.Tensor) -> torch.Tensor:
    """raw = ln(exp(6 sigma) - 1) (main_compression.py:51,58,64)."""
    return torch.log(torch.exp(sigma * 6) - 1)


def layer_param_counts(dims: Sequence[int]) -> List[int]:
    """Per-layer count out*(in+1) and nothing else (utils.py:215-231)."""
    return [dims[i + 1] * (dims[i] + 1) for i in range(len(dims) - 1)]


def gaussian_kl(q_loc, q_scale, p_loc, p_scale):
    """Closed-form KL(N(q)||N(p)) elementwise; same operation order as
    torch.distributions.kl._kl_normal_normal, which the reference calls at
    test_model.py:360,386 / prior_model.py:191-199,268."""
    var_ratio = (q_scale / p_scale) ** 2
    t1 = ((q_loc - p_loc) / p_scale) ** 2
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


# --------------------------------------------------------------------------- #
# modality description (the subset of config.py:28-137 the path needs)
# --------------------------------------------------------------------------- #
@dataclass
class Shape:
    dims: List[int]                       # [in, hidden..., out]
    data_dim: int
    pixel_sizes: List[int]
    upsample_factors: List[int]
    latent_dim: int = 128
    patch: bool = False
    patch_nums: Optional[List[int]] = None
    hier: Optional[Dict[str, List[int]]] = None   # hierarchical_patch_nums
    paddings: List[int] = field(default_factory=lambda: [2, 1, 1])
    layer_scales: List = field(default_factory=lambda: [4, 2, 2])
    w0: float = 30.0

    @property
    def n_weights(self) -> int:
        return int(sum(layer_param_counts(self.dims)))

    @property
    def lpe_dims(self) -> List[int]:
        return [self.pixel_sizes[i] // self.upsample_factors[i] for i in range(self.data_dim)]

    @property
    def n_latent(self) -> int:
        return int(np.prod(self.lpe_dims)) * self.latent_dim

    @property
    def n_pixels(self) -> int:
        return int(np.prod(self.pixel_sizes))


# --------------------------------------------------------------------------- #
# upsampling network (prior_model.py:23-59)
# --------------------------------------------------------------------------- #
def upsample_forward(w: Dict[str, torch.Tensor], x: torch.Tensor, shape: Shape) -> torch.Tensor:
    """nearest-up -> conv k5 -> lrelu -> up -> conv k3 -> lrelu -> up -> conv k3.

    `w` holds conv{1,2,3}.{weight,bias} in torch layout (oc, ic, *k).  x is
    (B, 128, *grid).  prior_model.py:47-59; LeakyReLU slope 0.01 (prior_model.py:42,44).
    """
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[shape.data_dim]
    for i in (1, 2, 3):
        f = shape.layer_scales[i - 1]
        f = tuple(float(v) for v in f) if isinstance(f, (tuple, list)) else float(f)
        x = F.interpolate(x, scale_factor=f, mode="nearest")
        x = conv(x, w[f"conv{i}.weight"], w[f"conv{i}.bias"], padding=shape.paddings[i - 1])
        if i < 3:
            x = F.leaky_relu(x, 0.01)
    return x


def latent_to_pe(w_up, lpe: torch.Tensor, shape: Shape) -> torch.Tensor:
    """(S, N, L) latent samples -> (N, S, pixels, 16) positional encodings.

    utils.py:4-120.  Non-patch: each row is upsampled on its own.  Patch: the
    rows of one datum are stitched into one grid (patch index and in-patch index
    interleaved per axis, utils.py:71-90), upsampled as a whole, and cut back
    into patches (utils.py:104-116).
    """
    S, N = lpe.shape[:2]
    g = shape.lpe_dims
    d = shape.data_dim
    C = shape.latent_dim
    z = lpe.reshape(S, N, *g, C)
    if not shape.patch:
        z = z.movedim(-1, 2).reshape(S * N, C, *g)
        pe = upsample_forward(w_up, z, shape)               # (S*N, 16, *pix)
        pe = pe.movedim(1, -1).reshape(S, N, -1, pe.shape[1])
    else:
        pn = shape.patch_nums
        z = z.reshape(S, -1, *pn, *g, C)
        # (S, data, p0, p1.., g0, g1.., C) -> (S, data, p0, g0, p1, g1, .., C)
        order = [0, 1]
        for i in range(d):
            order += [2 + i, 2 + d + i]
        z = z.permute(*order, 2 + 2 * d)
        z = z.reshape(S, -1, *[pn[i] * g[i] for i in range(d)], C)
        n_data = z.shape[1]
        z = z.movedim(-1, 2).reshape(S * n_data, C, *z.shape[2:-1])
        pe = upsample_forward(w_up, z, shape)               # (S*data, 16, *full)
        co = pe.shape[1]
        pe = pe.movedim(1, -1)
        split = []
        for i in range(d):
            split += [pn[i], shape.pixel_sizes[i]]
        pe = pe.reshape(S, n_data, *split, co)
        order = [0, 1] + [2 + 2 * i for i in range(d)] + [3 + 2 * i for i in range(d)] + [2 + 2 * d]
        pe = pe.permute(*order).reshape(S, N, -1, co)
    return pe.permute(1, 0, 2, 3)


# --------------------------------------------------------------------------- #
# hierarchical weight sampling (utils.py:122-198)
# --------------------------------------------------------------------------- #
def expand_level2(t: torch.Tensor, shape: Shape) -> torch.Tensor:
    """Repeat level-2 rows over the patches of their group (utils.py:151-180)."""
    d = shape.data_dim
    l2 = shape.hier["level2"]
    ng = [shape.patch_nums[i] // l2[i] for i in range(d)]
    W = t.shape[-1]
    t = t.reshape(-1, *ng, W)
    # insert a repeat axis after every group axis
    idx = [slice(None)]
    rep = [1]
    for i in range(d):
        idx += [slice(None), None]
        rep += [1, l2[i]]
    idx.append(slice(None))
    rep.append(1)
    t = t[tuple(idx)].repeat(rep)
    return t.reshape(-1, W)


def expand_level3(t: torch.Tensor, shape: Shape) -> torch.Tensor:
    """Repeat level-3 rows over all patches of the datum (utils.py:184-185)."""
    n = int(np.prod(shape.patch_nums))
    return t[:, None, :].repeat(1, n, 1).reshape(-1, t.shape[-1])


def sample_weights(loc, scale, eps_w, shape: Shape,
                   h_loc=None, h_scale=None, eps_h=None,
                   hh_loc=None, hh_scale=None, eps_hh=None) -> torch.Tensor:
    """h_w (N,S,W) = mu + sigma*eps  [+ level-2 + level-3 terms, each with its
    own per-patch noise] (utils.py:142-198).  eps_* are (N,S,W)."""
    hw = loc[:, None, :] + scale[:, None, :] * eps_w
    if shape.patch:
        hl, hs = expand_level2(h_loc, shape), expand_level2(h_scale, shape)
        hw = hw + (hl[:, None, :] + eps_h * hs[:, None, :])
        gl, gs = expand_level3(hh_loc, shape), expand_level3(hh_scale, shape)
        hw = hw + (gl[:, None, :] + gs[:, None, :] * eps_hh)
    return hw


# --------------------------------------------------------------------------- #
# INR forward (test_model.py:347-355, prior_model.py:168-179)
# --------------------------------------------------------------------------- #
def inr_forward(x_in: torch.Tensor, hw: torch.Tensor, A: Sequence[torch.Tensor], shape: Shape) -> torch.Tensor:
    """x_in (N,S,pix,in), hw (N,S,W) -> (N,S,pix,out).

    Per layer: vec = hw[seg] @ A_l; bias = vec[:out]; W = vec[out:].reshape(in,out)
    (test_model.py:260-280); x <- sin(w0*(x W + b)) except after the last layer.
    """
    counts = layer_param_counts(shape.dims)
    off = 0
    x = x_in
    n_layers = len(counts)
    for l, c in enumerate(counts):
        fin, fout = shape.dims[l], shape.dims[l + 1]
        vec = hw[..., off:off + c] @ A[l]
        off += c
        b = vec[..., :fout][:, :, None, :]
        Wm = vec[..., fout:].reshape(*vec.shape[:2], fin, fout)
        x = x @ Wm + b
        if l != n_layers - 1:
            x = torch.sin(shape.w0 * x)
    return x


# --------------------------------------------------------------------------- #
# test-time model state + predict (test_model.py:283-355)
# --------------------------------------------------------------------------- #
@dataclass
class Level:
    """One level of posterior state in *group order* (test_model.py:131-166,219-237)."""
    loc: torch.Tensor
    log_scale: torch.Tensor
    p_loc: torch.Tensor
    p_log_scale: torch.Tensor
    group_to_param: np.ndarray
    group_idx: np.ndarray
    group_start: np.ndarray
    group_end: np.ndarray
    mask: Optional[torch.Tensor] = None       # coded mask (param-wise)
    sample: Optional[torch.Tensor] = None     # coded values
    perm_g2p: Optional[np.ndarray] = None     # (rows, P) row permutation per column (patch only)

    def __post_init__(self):
        if self.mask is None:
            self.mask = torch.zeros_like(self.loc)
        if self.sample is None:
            self.sample = torch.zeros_like(self.loc)

    @property
    def n_groups(self) -> int:
        return len(self.group_start)

    def effective(self, permute_rows: bool = True):
        """mask-mix (test_model.py:289-290,320-321,327-328), un-permute columns
        (294-295,322-323) and go back to parameter order (297-298,324-325,329-330)."""
        m = self.mask
        loc = self.loc * (1 - m) + self.sample * m
        scale = std_transform(self.log_scale) * (1 - m) + 1e-15 * m
        if self.perm_g2p is not None and permute_rows:
            cols = torch.arange(loc.shape[1])[None, :].expand(loc.shape[0], -1)
            rows = torch.as_tensor(self.perm_g2p)
            loc, scale = loc[rows, cols], scale[rows, cols]
        g2p = torch.as_tensor(self.group_to_param)
        return loc[:, g2p], scale[:, g2p]


def column_row_permutations(n_rows: int, n_cols: int) -> np.ndarray:
    """Per-column row permutation, seeded by the column index (test_model.py:185-191)."""
    out = np.empty((n_rows, n_cols), dtype=np.int64)
    for c in range(n_cols):
        rs = np.random.RandomState(c)             # == np.random.seed(c); np.random.choice
        out[:, c] = rs.choice(n_rows, n_rows, False)
    return out


def predict(x, lvl1: Level, A, w_up, shape: Shape, eps: Dict[str, torch.Tensor], S: int,
            lvl2: Optional[Level] = None, lvl3: Optional[Level] = None) -> torch.Tensor:
    """One MC forward, returns (N,S,pix,out).  test_model.py:283-355.

    eps["lpe"] (S,N,L); eps["w"], eps["h"], eps["hh"] (N,S,W).
    """
    W = shape.n_weights
    loc, scale = lvl1.effective()
    lpe = loc[None, :, W:] + scale[None, :, W:] * eps["lpe"]
    pe = latent_to_pe(w_up, lpe, shape)
    xin = torch.cat([x[:, None].expand(-1, S, -1, -1), pe], -1)
    kw = {}
    if shape.patch:
        h_loc, h_scale = lvl2.effective()
        hh_loc, hh_scale = lvl3.effective(permute_rows=False)
        kw = dict(h_loc=h_loc, h_scale=h_scale, eps_h=eps["h"],
                  hh_loc=hh_loc, hh_scale=hh_scale, eps_hh=eps["hh"])
    hw = sample_weights(loc[:, :W], scale[:, :W], eps["w"], shape, **kw)
    return inr_forward(xin, hw, A, shape)


def weighted_kl(lvl: Level, beta: torch.Tensor) -> torch.Tensor:
    """sum_{n,p} beta[n,g(p)] KL[n,p]  (test_model.py:357-362)."""
    kl = gaussian_kl(lvl.loc, std_transform(lvl.log_scale),
                     lvl.p_loc[None, :], std_transform(lvl.p_log_scale)[None, :])
    return (kl * beta[:, torch.as_tensor(lvl.group_idx)]).sum()


def fit_loss(y_pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """mean over (n,s,pixel,channel) of squared error, times N (test_model.py:624-627)."""
    return torch.mean((y_pred - y[:, None]) ** 2) * y.shape[0]


def group_kl_nats(lvl: Level) -> np.ndarray:
    """Per-(row, block) KL in nats, f64 (test_model.py:384-388)."""
    with torch.no_grad():
        kl = gaussian_kl(lvl.loc, std_transform(lvl.log_scale),
                         lvl.p_loc[None, :], std_transform(lvl.p_log_scale)[None, :]).numpy()
    return np.stack([np.bincount(lvl.group_idx, weights=kl[i], minlength=lvl.n_groups)
                     for i in range(kl.shape[0])])


def anneal_beta(beta: torch.Tensor, kl_nats: np.ndarray, coded: np.ndarray,
                step: float = 0.05, upper: float = 0.0, lower: float = 0.4,
                bits: float = 16.0) -> torch.Tensor:
    """beta *= 1.05 where bits > 16+upper; /= 1.05 where bits <= 16-lower; clamp
    [0,1e4]; coded blocks unchanged (test_model.py:404-413)."""
    b = kl_nats / LN2
    up = torch.from_numpy(1 + step * (b > bits + upper).astype(float)).float()
    dn = torch.from_numpy(1 + step * (b <= bits - lower).astype(float)).float()
    nb = torch.clamp(beta * up / dn, 0.0, 10000.0)
    return torch.where(torch.from_numpy(~coded.astype(bool)), nb, beta)


# --------------------------------------------------------------------------- #
# REC (test_model.py:441-533)
# --------------------------------------------------------------------------- #
def gumbel_sequence(seed: int, n: int = 65536) -> np.ndarray:
    """Decreasing truncated-Gumbel sequence shared by all blocks (test_model.py:441-457):
    g_0 = -ln(-ln u_0); g_i = -ln(-ln u_i + exp(-g_{i-1})), u from legacy MT19937."""
    neg_log_u = -np.log(np.random.RandomState(seed).rand(n))
    g = np.empty(n)
    acc = 0.0                       # running exp(-g_{i-1}); 0 for the first point
    for i in range(n):
        acc = neg_log_u[i] + acc
        g[i] = -np.log(acc)
        acc = np.exp(-g[i])         # re-derive from the rounded g_i, as the reference does
    return g


def sobol_words(D: int, seed: int):
    """Scrambled-Sobol state of torch's engine: shift (D,), direction words (D,30)."""
    eng = torch.quasirandom.SobolEngine(D, scramble=True, seed=seed)
    return eng.shift.numpy().copy(), eng.sobolstate.numpy().copy()


def sobol_uniform(D: int, n: int, seed: int) -> np.ndarray:
    """First n points as f32 (SobolEngine.draw; test_model.py:494-495), by random access:
    q_k = shift ^ XOR_{j in bits(gray(k))} word[:, j]; u = f32(q) * 2^-30."""
    shift, words = sobol_words(D, seed)
    k = np.arange(n, dtype=np.int64)
    gray = k ^ (k >> 1)
    q = np.broadcast_to(shift[None, :], (n, D)).copy()
    for j in range(int(max(1, n - 1)).bit_length()):
        sel = ((gray >> j) & 1).astype(bool)
        q[sel] ^= words[:, j][None, :]
    return q.astype(np.float32) * np.float32(2.0 ** -30)


_P0 = [-5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1,
       1.39312609387279679503E1, -1.23916583867381258016E0]
_Q0 = [1.95448858338141759834E0, 4.67627912898881538453E0, 8.63602421390890590575E1,
       -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1,
       1.59056225126211695515E1, -1.18331621121330003142E0]
_P1 = [4.05544892305962419923E0, 3.15251094599893866154E1, 5.71628192246421288162E1,
       4.40805073893200834700E1, 1.46849561928858024014E1, 2.18663306850790267539E0,
       -1.40256079171354495875E-1, -3.50424626827848203418E-2, -8.57456785154685413611E-4]
_Q1 = [1.57799883256466749731E1, 4.53907635128879210584E1, 4.13172038254672030440E1,
       1.50425385692907503408E1, 2.50464946208309415979E0, -1.42182922854787788574E-1,
       -3.80806407691578277194E-2, -9.33259480895457427372E-4]
_P2 = [3.23774891776946035970E0, 6.91522889068984211695E0, 3.93881025292474443415E0,
       1.33303460815807542389E0, 2.01485389549179081538E-1, 1.23716634817820021358E-2,
       3.01581553508235416007E-4, 2.65806974686737550832E-6, 6.23974539184983293730E-9]
_Q2 = [6.02427039364742014255E0, 3.67983563856160859403E0, 1.37702099489081330271E0,
       2.16236993594496635890E-1, 1.34204006088543189037E-2, 3.28014464682127739104E-4,
       2.89247864745380683936E-6, 6.79019408009981274425E-9]


def _horner(x, coef, monic=False):
    r = (x + coef[0]) if monic else np.full_like(x, coef[0])
    for c in coef[1:]:
        r = r * x + c
    return r


def ndtri_cephes(p: np.ndarray) -> np.ndarray:
    """Inverse normal CDF, f64.  Third-party arithmetic on the path: the reference
    calls scipy.stats.norm.ppf (test_model.py:496), i.e. Cephes `ndtri` (scipy is
    unpinned in requirements.txt; 1.18.1 installed here).  This restates the
    published Cephes algorithm (three rational approximations); it agrees with
    scipy.special.ndtri to <= 8e-16 relative (tests/test_oracle_golden.py)."""
    p = np.asarray(p, np.float64)
    out = np.empty_like(p)
    tail = p.copy()
    flip = tail > 1.0 - 0.13533528323661269189
    tail[flip] = 1.0 - tail[flip]
    central = tail > 0.13533528323661269189
    y = tail[central] - 0.5
    y2 = y * y
    out[central] = (y + y * (y2 * _horner(y2, _P0) / _horner(y2, _Q0, True))) * 2.50662827463100050242
    t = tail[~central]
    with np.errstate(divide="ignore", invalid="ignore"):
        x = np.sqrt(-2.0 * np.log(t))
        x0 = x - np.log(x) / x
        z = 1.0 / x
        x1 = np.where(x < 8.0,
                      z * _horner(z, _P1) / _horner(z, _Q1, True),
                      z * _horner(z, _P2) / _horner(z, _Q2, True))
        r = x0 - x1
    r = np.where(flip[~central], r, -r)
    out[~central] = r
    out[p == 0.0] = -np.inf
    out[p == 1.0] = np.inf
    return out


def candidate_table(D: int, n: int, seed: int) -> np.ndarray:
    """Standard-normal candidates (n, D) as the reference builds them
    (test_model.py:493-498): Sobol f32 -> norm.ppf -> clamp +-100.  scipy's ppf on an
    f32 array runs its f->f loop, so every entry is an f32 value stored in f64
    (verified in make_golden.py: table == f32(ndtri_f64(u)))."""
    u = sobol_uniform(D, n, seed)
    s = ndtri_cephes(u.astype(np.float64)).astype(np.float32).astype(np.float64)
    return np.clip(s, -100.0, 100.0)


def rec_log_weights(q_loc, q_scale, p_loc, p_scale, table: np.ndarray, gumbel: np.ndarray) -> np.ndarray:
    """log_w_k = sum_d log q(z_kd) - log p(z_kd) + g_k in f64 with the reference's
    mixed-precision steps (test_model.py:512-526; torch Normal.log_prob):
    scale**2 and log(scale) are rounded to f32 before entering the f64 expression."""
    f32 = np.float32
    q_loc, q_scale, p_loc, p_scale = (np.asarray(a, f32) for a in (q_loc, q_scale, p_loc, p_scale))
    z = p_loc.astype(np.float64) + p_scale.astype(np.float64) * table
    half_log_2pi = math.log(math.sqrt(2 * math.pi))

    def log_prob(loc, scale):
        var = (scale * scale).astype(np.float64)          # f32 square, then promoted
        # f32 log via torch (the reference's library: np.log differs from it by 1 f32 ulp
        # on some inputs); a per-block constant, so it never affects the argmax
        log_scale = torch.from_numpy(scale).log().numpy().astype(np.float64)
        return -((z - loc.astype(np.float64)) ** 2) / (2 * var) - log_scale - half_log_2pi

    lw = log_prob(q_loc, q_scale).sum(-1) - log_prob(p_loc, p_scale).sum(-1)
    return lw + gumbel[: table.shape[0]]


def rec_encode(q_loc, q_scale, p_loc, p_scale, table, gumbel):
    """(index, z_i as f32, log_w) -- first-max argmax, sample truncated to f32 on
    store (test_model.py:530-531,591)."""
    lw = rec_log_weights(q_loc, q_scale, p_loc, p_scale, table, gumbel)
    i = int(np.argmax(lw))
    z = np.asarray(p_loc, np.float32).astype(np.float64) + np.asarray(p_scale, np.float32).astype(np.float64) * table[i]
    return i, z.astype(np.float32), lw


def rec_decode(index: int, p_loc, p_scale, table) -> np.ndarray:
    """Regenerate the coded sample from its index (the reference has no decoder;
    this is the receiver-side half of test_model.py:514,531)."""
    z = np.asarray(p_loc, np.float32).astype(np.float64) + np.asarray(p_scale, np.float32).astype(np.float64) * table[index]
    return z.astype(np.float32)


# --------------------------------------------------------------------------- #
# block grouping (prior_model.py:264-316)
# --------------------------------------------------------------------------- #
def grouping_by_kl(bits: np.ndarray, cap: float = 16.0):
    """Seed-0 permutation then greedy sequential binning with per-block cap
    (prior_model.py:273-316).  Returns the reference's 8-tuple."""
    P = bits.shape[0]
    order = np.random.RandomState(0).choice(P, P, False)
    w = bits[order]
    starts = [0]
    acc = w[0]
    for i in range(1, P):
        if acc + w[i] > cap:
            starts.append(i)
            acc = w[i]
        else:
            acc = acc + w[i]
    starts = np.array(starts)
    ends = np.append(starts[1:], P)
    n_groups = len(starts)
    param2group = order.copy()
    group2param = np.argsort(param2group)
    group_idx = np.repeat(np.arange(n_groups), ends - starts).astype(int)
    # python-float sequential sums, as the reference's sum([...]) does
    group_kls = np.array([sum([bits[i] for i in order[s:e]]) for s, e in zip(starts, ends)])
    return group_idx, starts, ends, group2param, param2group, n_groups, group_kls, bits


# --------------------------------------------------------------------------- #
# prior training pieces (prior_model.py:129-200, main_prior_training.py:135-172)
# --------------------------------------------------------------------------- #
def prior_forward(x, loc, log_scale, lpe_loc, lpe_log_scale, A, w_up, shape: Shape, eps,
                  h_loc=None, h_log_scale=None, hh_loc=None, hh_log_scale=None):
    """Single-sample forward in parameter order (prior_model.py:129-179).
    eps["lpe"] has lpe_loc's shape; eps["w"|"h"|"hh"] are (N,1,W)."""
    lpe = lpe_loc + std_transform(lpe_log_scale) * eps["lpe"]
    pe = latent_to_pe(w_up, lpe.reshape(1, lpe.shape[0], -1), shape)[:, 0]
    kw = {}
    if shape.patch:
        kw = dict(h_loc=h_loc, h_scale=std_transform(h_log_scale), eps_h=eps["h"],
                  hh_loc=hh_loc, hh_scale=std_transform(hh_log_scale), eps_hh=eps["hh"])
    hw = sample_weights(loc, std_transform(log_scale), eps["w"], shape, **kw)
    xin = torch.cat([x, pe], -1)[:, None]
    return inr_forward(xin, hw, A, shape)[:, 0]


def em_prior_update(loc: torch.Tensor, log_scale: torch.Tensor):
    """mu_p = mean_n mu_q; sigma_p = sqrt(mean_n sigma_q^2 + var_n mu_q), unbiased
    variance (main_prior_training.py:157-159)."""
    mu = loc.mean(0)
    var = (std_transform(log_scale) ** 2).mean(0) + loc.var(0)
    return mu, var ** 0.5


def beta_controller(beta: float, kl_bits: float, budget_min: float, budget_max: float) -> float:
    """x1.5 above budget, /1.5 below, clamp [1e-20, 1] (main_prior_training.py:146-154)."""
    if kl_bits > budget_max:
        beta *= 1.5
    if kl_bits < budget_min:
        beta /= 1.5
    return min(max(beta, 1e-20), 1)


# --------------------------------------------------------------------------- #
# inputs (utils.py:265-298 + data/image.py:24-27)
# --------------------------------------------------------------------------- #
def fourier_inputs(sizes: Sequence[int], fourier_dim: int) -> torch.Tensor:
    """Pixel-centre grid in (-1,1) and [cos(pi x w), sin(pi x w)] features with
    w = exp(linspace(0, ln 1024, fourier_dim/(2 d)))  -> (pixels, fourier_dim)."""
    d = len(sizes)
    axes = [-1 + 2 * (0.5 + torch.arange(s)) / s for s in sizes]
    grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), -1).reshape(-1, d)
    w = torch.exp(torch.linspace(0, math.log(1024), fourier_dim // (2 * d)))
    arg = (grid[:, :, None] * w[None, None, :]).reshape(grid.shape[0], -1)
    return torch.cat([torch.cos(math.pi * arg), torch.sin(math.pi * arg)], -1)
