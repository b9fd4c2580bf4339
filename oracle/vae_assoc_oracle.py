"""fp64 numpy restatement of the reference's associated-VAE train step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED by the
reference (it has no tests / golden vectors and TensorFlow is not installable
here).  Every function cites the reference lines (/root/reference/...) it
follows.  The backward pass is derived by hand; tests/test_oracle_grads.py
checks it against torch autograd on an independent restatement of the same
graph (oracle/torch_twin.py) and against central finite differences.

Parameter order per modality = the reference's tf.Variable creation order
(vae_assoc.py:185-215 encoder, :257-300 decoder):

  dense modality (hidden_conv=False), 14 tensors
    0 W1 [n_input, r1]   1 b1 [r1]    2 W2 [r1, r2]   3 b2 [r2]
    4 Wmu [r2, n_z]      5 bmu [n_z]  6 Wls [r2, n_z] 7 bls [n_z]
    8 V1 [n_z, r1]       9 c1 [r1]   10 V2 [r1, r2]  11 c2 [r2]     <- decoder uses the RECOG sizes
   12 Vo [r2, n_input]  13 co [n_input]                                 (vae_assoc.py:257,280,293,299)

  conv modality (hidden_conv=True), 17 tensors
    0 C1 [5,5,1,r1]  1 C2 [5,5,r1,2r1]  2 C3 [5,5,2r1,r2]               (vae_assoc.py:174-197, no bias, no act)
    3 Wmu [s*s*r2, n_z] 4 bmu 5 Wls 6 bls                               (vae_assoc.py:207-210)
    7 D1 [3,3,g1,n_z] 8 d1 [g1]   9 D2 [5,5,g1/2,g1] 10 d2 [g1/2]       (vae_assoc.py:251-277 -> deconv.py:78,110-114)
   11 D3 [5,5,g2,g1/2] 12 d3 [g2] 13 D4 [5,5,1,g2] 14 d4 [1]
   15 Wo [n_input, n_input] 16 bo [n_input]                             (vae_assoc.py:287-291)
"""
import itertools

import numpy as np

ADAM_BETA1 = 0.9      # tf.train.AdamOptimizer defaults (vae_assoc.py:373-374; TensorFlow, not vendored)
ADAM_BETA2 = 0.999
ADAM_EPS = 1e-8
CE_EPS = 1e-3         # vae_assoc.py:322-323 (the comment says 1e-10, the code says 1e-3)


# --------------------------------------------------------------------------------------
# initialisation (vae_assoc.py:11-18, :471-473, deconv.py:81-84)
# --------------------------------------------------------------------------------------
def xavier_init(rng, fan_in, fan_out, constant=1.0):
    """vae_assoc.py:11-18: U(-c*sqrt(6/(fi+fo)), +c*sqrt(6/(fi+fo))), shape [fan_in, fan_out]."""
    lim = constant * np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out))


def truncated_normal(rng, shape, stddev=0.1):
    """vae_assoc.py:471-473 weight_variable: tf.truncated_normal re-draws samples beyond 2 sigma."""
    out = rng.normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return out * stddev


def deconv_xavier(rng, kh, kw, out_depth, in_depth):
    """deconv.py:78-84: filter [kh,kw,out_depth,in_depth]; prettytensor layers.xavier_init(out*k*k, in*k*k)
    (uniform, limit sqrt(6/(n_in+n_out)); prettytensor is third-party and unpinned)."""
    patch = kh * kw
    lim = np.sqrt(6.0 / (out_depth * patch + in_depth * patch))
    return rng.uniform(-lim, lim, size=(kh, kw, out_depth, in_depth))


def conv_geometry(na):
    """Sizes of the conv variant for one architecture dict (vae_assoc.py:171,192,198)."""
    s0 = int(round(np.sqrt(na["n_input"])))
    assert s0 * s0 == na["n_input"]
    s1 = (s0 + 1) // 2           # SAME, stride 2
    s2 = (s1 + 1) // 2           # SAME, stride 2  (reference: input_size/2/2, exact for 28)
    s3 = s2 - 5 + 1              # VALID, stride 1
    return s0, s1, s2, s3


def param_names(na):
    if na["hidden_conv"]:
        return ["C1", "C2", "C3", "Wmu", "bmu", "Wls", "bls",
                "D1", "d1", "D2", "d2", "D3", "d3", "D4", "d4", "Wo", "bo"]
    return ["W1", "b1", "W2", "b2", "Wmu", "bmu", "Wls", "bls", "V1", "c1", "V2", "c2", "Vo", "co"]


def param_shapes(na):
    nz, ni = na["n_z"], na["n_input"]
    r1, r2 = na["n_hidden_recog_1"], na["n_hidden_recog_2"]
    if not na["hidden_conv"]:
        return [(ni, r1), (r1,), (r1, r2), (r2,), (r2, nz), (nz,), (r2, nz), (nz,),
                (nz, r1), (r1,), (r1, r2), (r2,), (r2, ni), (ni,)]
    g1, g2 = na["n_hidden_gener_1"], na["n_hidden_gener_2"]
    _, _, _, s3 = conv_geometry(na)
    flat = s3 * s3 * r2
    return [(5, 5, 1, r1), (5, 5, r1, 2 * r1), (5, 5, 2 * r1, r2),
            (flat, nz), (nz,), (flat, nz), (nz,),
            (3, 3, g1, nz), (g1,), (5, 5, g1 // 2, g1), (g1 // 2,),
            (5, 5, g2, g1 // 2), (g2,), (5, 5, 1, g2), (1,),
            (ni, ni), (ni,)]


def init_params(archs, seed=0):
    """Reference initialisers, drawn from one numpy RandomState (the reference's TF stream is not reproducible)."""
    rng = np.random.RandomState(seed)
    out = []
    for na in archs:
        ps = []
        for name, shp in zip(param_names(na), param_shapes(na)):
            if len(shp) == 1:
                ps.append(np.zeros(shp))                      # all biases zero (vae_assoc.py:186 ..., deconv.py:113)
            elif name in ("C1", "C2", "C3"):
                ps.append(truncated_normal(rng, shp))
            elif name in ("D1", "D2", "D3", "D4"):
                ps.append(deconv_xavier(rng, *shp))
            else:
                ps.append(xavier_init(rng, *shp))
        out.append(ps)
    return out


# --------------------------------------------------------------------------------------
# elementwise pieces
# --------------------------------------------------------------------------------------
def act(name, a):
    if name == "relu":
        return np.maximum(a, 0.0)
    if name == "softplus":
        return np.logaddexp(0.0, a)
    raise ValueError(name)


def act_grad_from_output(name, h):
    """d act / d pre-activation expressed through the OUTPUT h (what the CUDA path stores):
    relu: 1[h>0] (TF ReluGrad uses y>0);  softplus: sigmoid(a) = 1 - exp(-h)."""
    if name == "relu":
        return (h > 0).astype(h.dtype)
    if name == "softplus":
        return 1.0 - np.exp(-h)
    raise ValueError(name)


def sigmoid(a):
    return 0.5 * (1.0 + np.tanh(0.5 * a))


def tf32_round(x):
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) to a 10-bit mantissa, kept in an fp32 container."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + np.uint64(0x1000)) & np.uint64(0xFFFFE000)).astype(np.uint32)
    return u.view(np.float32).astype(np.float64)


def tc_served(batch, aligned=True):
    """Mirror of tc_supported() in vae_assoc_b200/csrc/gemm_tc2.cu: a contraction runs on tcgen05 kind::tf32 iff its
    batch extent is >= 32 and every row pitch is a multiple of 4 floats (the library pads all pitches except those of
    the [B, n_z] / [B, 2 n_z] latent tensors)."""
    return batch >= 32 and aligned


def tf32_plan(na, B):
    """Where the CUDA tf32 path rounds (RNA) operands, for one dense modality at batch B -- the same producer-side
    rules as build_ops() in vae_assoc_b200/csrc/api.cu: a tensor is rounded by its producer iff one of its consumers
    is a tensor-core GEMM; tensor-core GEMMs read the tf32-rounded shadow of the weights, SIMT ones the master."""
    ni, r1, r2, nz = na["n_input"], na["n_hidden_recog_1"], na["n_hidden_recog_2"], na["n_z"]
    a4, a2 = nz % 4 == 0, nz % 2 == 0          # pitch of z / dz is n_z, of the heads / dheads 2 n_z
    t = dict(
        f_e1=tc_served(B), f_e2=tc_served(B), f_hd=tc_served(B, a2),
        f_d1=tc_served(B, a4), f_d2=tc_served(B), f_o=tc_served(B),
        w_o=tc_served(B), d_o=tc_served(B), w_d2=tc_served(B), d_d2=tc_served(B),
        w_d1=tc_served(B, a4), d_d1=tc_served(B, a4), w_hd=tc_served(B, a2), d_hd=tc_served(B, a2),
        w_e2=tc_served(B), d_e2=tc_served(B), w_e1=tc_served(B))
    t.update(r_x=t["f_e1"] or t["w_e1"], r_h1=t["f_e2"] or t["w_e2"], r_h2=t["f_hd"] or t["w_hd"],
             r_z=t["f_d1"] or t["w_d1"], r_g1=t["f_d2"] or t["w_d2"], r_g2=t["f_o"] or t["w_o"],
             r_da=t["w_o"] or t["d_o"], r_dg2=t["w_d2"] or t["d_d2"], r_dg1=t["w_d1"] or t["d_d1"],
             r_dhd=t["w_hd"] or t["d_hd"], r_dh2=t["w_e2"] or t["d_e2"], r_dh1=t["w_e1"])
    return t


# --------------------------------------------------------------------------------------
# TF conv2d / conv2d_transpose (NHWC, filter [kh,kw,cin,cout]); semantics from TensorFlow's documented
# padding rules (third-party, not vendored): SAME: out = ceil(in/s), pad_total = max((out-1)s+k-in,0),
# pad_before = pad_total // 2;  VALID: out = (in-k)//s + 1, no padding.
# --------------------------------------------------------------------------------------
def conv_out_size(n, k, s, padding):
    if padding == "SAME":
        o = -(-n // s)
        pad_total = max((o - 1) * s + k - n, 0)
        return o, pad_total // 2
    return (n - k) // s + 1, 0


def _patches(x, k, s, pad_before, out):
    """x [B,H,W,C] -> [B,out,out,k,k,C] (zero padded)."""
    B, H, W, C = x.shape
    need = (out - 1) * s + k
    xp = np.zeros((B, max(need, H + pad_before), max(need, W + pad_before), C), dtype=x.dtype)
    xp[:, pad_before:pad_before + H, pad_before:pad_before + W, :] = x
    p = np.empty((B, out, out, k, k, C), dtype=x.dtype)
    for ky in range(k):
        for kx in range(k):
            p[:, :, :, ky, kx, :] = xp[:, ky:ky + (out - 1) * s + 1:s, kx:kx + (out - 1) * s + 1:s, :]
    return p


def _col2im(p, H, W, k, s, pad_before):
    """adjoint of _patches: p [B,out,out,k,k,C] -> [B,H,W,C]."""
    B, out = p.shape[0], p.shape[1]
    C = p.shape[5]
    need = (out - 1) * s + k
    xp = np.zeros((B, max(need, H + pad_before), max(need, W + pad_before), C), dtype=p.dtype)
    for ky in range(k):
        for kx in range(k):
            xp[:, ky:ky + (out - 1) * s + 1:s, kx:kx + (out - 1) * s + 1:s, :] += p[:, :, :, ky, kx, :]
    return xp[:, pad_before:pad_before + H, pad_before:pad_before + W, :]


def conv2d(x, w, s, padding):
    """tf.nn.conv2d (vae_assoc.py:484-486)."""
    k = w.shape[0]
    out, pb = conv_out_size(x.shape[1], k, s, padding)
    return np.tensordot(_patches(x, k, s, pb, out), w, axes=([3, 4, 5], [0, 1, 2]))


def conv2d_bwd(x, w, s, padding, dy):
    k = w.shape[0]
    out, pb = conv_out_size(x.shape[1], k, s, padding)
    p = _patches(x, k, s, pb, out)
    dw = np.tensordot(p, dy, axes=([0, 1, 2], [0, 1, 2]))
    dp = np.tensordot(dy, w, axes=([3], [3]))                  # [B,out,out,k,k,cin]
    dx = _col2im(dp, x.shape[1], x.shape[2], k, s, pb)
    return dx, dw


def deconv_out_size(n, k, s, padding):
    """deconv.py:133-161 get2d_deconv_output_size."""
    return (n - 1) * s + k if padding == "VALID" else n * s


def conv2d_transpose(y, w, s, padding):
    """tf.nn.conv2d_transpose (deconv.py:107) = gradient of conv2d wrt its input.
    y [B,h,h,in_depth], w [k,k,out_depth,in_depth] -> [B,H,H,out_depth]."""
    k = w.shape[0]
    H = deconv_out_size(y.shape[1], k, s, padding)
    o, pb = conv_out_size(H, k, s, padding)
    assert o == y.shape[1]
    dp = np.tensordot(y, w, axes=([3], [3]))                   # [B,h,h,k,k,out_depth]
    return _col2im(dp, H, H, k, s, pb)


def conv2d_transpose_bwd(y, w, s, padding, dout):
    """returns (dy, dw) for out = conv2d_transpose(y, w)."""
    k = w.shape[0]
    H = dout.shape[1]
    o, pb = conv_out_size(H, k, s, padding)
    p = _patches(dout, k, s, pb, o)                            # [B,h,h,k,k,out_depth]
    dy = np.tensordot(p, w, axes=([3, 4, 5], [0, 1, 2]))       # [B,h,h,in_depth]
    dw = np.tensordot(p, y, axes=([0, 1, 2], [0, 1, 2]))       # [k,k,out_depth,in_depth]
    return dy, dw


# --------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------
class _Never(dict):
    def __missing__(self, key):
        return False


_NO_ROUNDING = _Never()


class OracleAssocVAE(object):
    """Same constructor surface as the reference class (vae_assoc.py:26-27), numpy fp64 arithmetic."""

    def __init__(self, network_architectures, binary=True, transfer_fct="softplus", weights=1.0,
                 assoc_lambda=1.0, learning_rate=0.001, batch_size=100, params=None, seed=0,
                 dtype=np.float64, emulate_tf32=False):
        self.network_architectures = network_architectures
        self.assoc_lambda = assoc_lambda
        n = len(network_architectures)
        self.binary = list(binary) if isinstance(binary, (list, tuple)) else [binary] * n       # :31-35
        self.weights = list(weights) if isinstance(weights, (list, tuple)) else [weights] * n   # :37-41
        assert len(self.binary) == n and len(self.weights) == n
        self.transfer_fct = getattr(transfer_fct, "__name__", transfer_fct)
        self.learning_rate = learning_rate
        self.batch_size = batch_size
        self.n_z = network_architectures[0]["n_z"]                                               # :89
        self.dtype = dtype
        # emulate_tf32: round operands to tf32 exactly where the CUDA tensor-core path does (dense modalities);
        # products of tf32 numbers are exact in fp32, so this oracle then differs from the GPU only by fp32
        # accumulation order -- it pins the tf32 path to ~1e-5 instead of the 2e-3 "tf32 vs exact" bound.
        self.emulate_tf32 = emulate_tf32
        if params is None:
            params = init_params(network_architectures, seed)
        self.params = [[np.array(p, dtype=dtype) for p in ps] for ps in params]
        self.m = [[np.zeros_like(p) for p in ps] for ps in self.params]
        self.v = [[np.zeros_like(p) for p in ps] for ps in self.params]
        self.t = 0
        self.rng = np.random.RandomState(seed + 1)

    # ---- tf32 emulation helpers ------------------------------------------------------
    def _plan(self, m, B):
        na = self.network_architectures[m]
        if not self.emulate_tf32 or na["hidden_conv"]:
            return _NO_ROUNDING
        return tf32_plan(na, B)

    @staticmethod
    def _q(flag, a):
        return tf32_round(a) if flag else a

    # ---- forward pieces -------------------------------------------------------------
    def encode(self, m, x):
        """_recognition_network (vae_assoc.py:163-222).  Returns (mu, logvar, cache)."""
        na, P, f = self.network_architectures[m], self.params[m], self.transfer_fct
        if na["hidden_conv"]:
            s0 = conv_geometry(na)[0]
            x2 = x.reshape(-1, s0, s0, 1)                                                        # :172
            l05 = conv2d(x2, P[0], 2, "SAME")                                                    # :174-178 (no act)
            l1 = conv2d(l05, P[1], 2, "SAME")                                                    # :179-183
            l2 = conv2d(l1, P[2], 1, "VALID")                                                    # :194-197
            h2 = l2.reshape(l2.shape[0], -1)                                                     # :199
            mu = h2 @ P[3] + P[4]                                                                # :217-218
            lv = h2 @ P[5] + P[6]                                                                # :219-221
            return mu, lv, dict(x2=x2, l05=l05, l1=l1, l2shape=l2.shape, h2=h2)
        t = self._plan(m, x.shape[0])
        x = self._q(t["r_x"], x)
        h1 = self._q(t["r_h1"], act(f, x @ self._q(t["f_e1"], P[0]) + P[1]))                     # :187-188
        h2 = self._q(t["r_h2"], act(f, h1 @ self._q(t["f_e2"], P[2]) + P[3]))                    # :203-204
        mu = h2 @ self._q(t["f_hd"], P[4]) + P[5]
        lv = h2 @ self._q(t["f_hd"], P[6]) + P[7]
        return mu, lv, dict(x=x, h1=h1, h2=h2)

    def decode(self, m, z):
        """_generator_network (vae_assoc.py:243-304).  Returns (x_reconstr_mean, cache)."""
        na, P, f = self.network_architectures[m], self.params[m], self.transfer_fct
        if na["hidden_conv"]:
            if not self.binary[m]:
                raise ValueError("hidden_conv with a Gaussian output is ill-formed in the reference "
                                 "(vae_assoc.py:299-303 multiplies [B,n_input] by [recog_2,n_input])")
            z2 = z.reshape(-1, 1, 1, self.n_z)                                                   # :250
            # every deconv_2d call keeps the wrapper's default sigmoid (vae_assoc.py:491) -- SURVEY 3.3 trap
            a1 = conv2d_transpose(z2, P[7], 1, "VALID") + P[8];   o1 = sigmoid(a1)               # :251-255
            a2 = conv2d_transpose(o1, P[9], 1, "VALID") + P[10];  o2 = sigmoid(a2)               # :263-267
            a3 = conv2d_transpose(o2, P[11], 2, "SAME") + P[12];  o3 = sigmoid(a3)               # :268-272
            a4 = conv2d_transpose(o3, P[13], 2, "SAME") + P[14];  o4 = sigmoid(a4)               # :273-277
            g2 = o4.reshape(o4.shape[0], -1)                                                     # :278
            a = g2 @ P[15] + P[16]                                                               # :287-291
            return sigmoid(a), dict(z2=z2, o1=o1, o2=o2, o3=o3, o4=o4, g2=g2)
        t = self._plan(m, z.shape[0])
        g1 = self._q(t["r_g1"], act(f, z @ self._q(t["f_d1"], P[8]) + P[9]))                     # :257-260
        g2 = self._q(t["r_g2"], act(f, g1 @ self._q(t["f_d2"], P[10]) + P[11]))                  # :280-283
        a = g2 @ self._q(t["f_o"], P[12]) + P[13]
        xh = sigmoid(a) if self.binary[m] else a                                                 # :285-303
        return xh, dict(z=z, g1=g1, g2=g2)

    def forward(self, X, eps):
        """_create_network (vae_assoc.py:78-119): ONE eps [B,n_z] shared by all modalities (:90)."""
        out = dict(z_means=[], z_log_sigma_sqs=[], z_array=[], x_reconstr_means=[], enc=[], dec=[])
        for m in range(len(self.network_architectures)):
            mu, lv, ec = self.encode(m, np.asarray(X[m], dtype=self.dtype))
            z = mu + np.sqrt(np.exp(lv)) * eps                                                   # :102-103
            z = self._q(self._plan(m, z.shape[0])["r_z"], z)
            xh, dc = self.decode(m, z)
            out["z_means"].append(mu); out["z_log_sigma_sqs"].append(lv); out["z_array"].append(z)
            out["x_reconstr_means"].append(xh); out["enc"].append(ec); out["dec"].append(dc)
        return out

    # ---- loss (vae_assoc.py:306-371) ---------------------------------------------------
    def loss(self, X, fw, global_batch=None):
        """Returns cost and the probe quantities.  `global_batch` (default: local batch) is the divisor of the
        batch-MEAN terms; batch-SUM terms (l2_loss :328, assoc KL :355-365) are never divided, so that summing
        the per-shard costs of a batch-sharded run gives the single-process cost."""
        M = len(self.network_architectures)
        B = np.asarray(X[0]).shape[0]
        Bg = float(global_batch if global_batch is not None else B)
        rec, lat, costs = [], [], []
        for m in range(M):
            x = np.asarray(X[m], dtype=self.dtype)
            x = self._q(self._plan(m, x.shape[0])["r_x"], x)      # the loss kernel reads the staged copy of x
            xh, mu, lv = fw["x_reconstr_means"][m], fw["z_means"][m], fw["z_log_sigma_sqs"][m]
            if self.binary[m]:
                r = -np.sum(x * np.log(CE_EPS + xh) + (1 - x) * np.log(CE_EPS + 1 - xh), 1)      # :321-324  [B]
                r_sum = r.sum() / Bg
            else:
                r = 0.5 * np.sum((x - xh) ** 2)                                                  # :327-328  scalar
                # reduce_mean(scalar + [B]) = scalar + mean  (:340); under sharding each shard adds its part
                r_sum = r
            l = -0.5 * np.sum(1 + lv - mu ** 2 - np.exp(lv), 1)                                  # :335-337  [B]
            rec.append(r); lat.append(l)
            costs.append((r_sum + l.sum() / Bg) * self.weights[m])                               # :340
        assoc = []
        nz = self.n_z
        for i, j in itertools.combinations(range(M), 2):                                         # :346
            mi, mj = fw["z_means"][i], fw["z_means"][j]
            li, lj = fw["z_log_sigma_sqs"][i], fw["z_log_sigma_sqs"][j]
            a = np.sum(0.5 * (lj.sum(1) - li.sum(1) - nz + np.exp(li - lj).sum(1)
                              + ((mj - mi) ** 2 * np.exp(-lj)).sum(1)))                          # :355-359
            a += np.sum(0.5 * (li.sum(1) - lj.sum(1) - nz + np.exp(lj - li).sum(1)
                               + ((mi - mj) ** 2 * np.exp(-li)).sum(1)))                         # :361-365
            assoc.append(a)
        cost = sum(costs) + self.assoc_lambda * sum(assoc) if assoc else sum(costs)              # :368-371
        return dict(cost=cost, vae_costs=costs, vae_reconstr_losses=rec, vae_latent_losses=lat, assoc_costs=assoc)

    # ---- backward (what tf.train.AdamOptimizer.minimize differentiates, vae_assoc.py:373-374) ----
    def loss_and_grads(self, X, eps, global_batch=None, masks=None):
        """`masks` (optional, dense relu modalities): per modality a dict {"h1","h2","g1","g2"} of boolean [B, width]
        arrays that REPLACE relu'(.) = 1[h > 0] in the backward pass (TF ReluGrad, y > 0).  The tensor-core parity test
        reads the masks the CUDA path actually applied (vaeassoc_probe_mask) and passes them here, so that gradients
        are compared under identical masks and the residual is the kernels' arithmetic alone."""
        M = len(self.network_architectures)
        X = [np.asarray(x, dtype=self.dtype) for x in X]
        eps = np.asarray(eps, dtype=self.dtype)
        B = X[0].shape[0]
        Bg = float(global_batch if global_batch is not None else B)
        fw = self.forward(X, eps)
        ls = self.loss(X, fw, global_batch)
        f, lam = self.transfer_fct, self.assoc_lambda

        def ag(m, name, h):
            if masks is not None and masks[m] is not None:
                return np.asarray(masks[m][name], dtype=h.dtype)
            return act_grad_from_output(f, h)
        grads = [[None] * len(ps) for ps in self.params]
        dz_list = []
        # decoders
        for m in range(M):
            P, G, w = self.params[m], grads[m], self.weights[m]
            t = self._plan(m, B)
            x, xh, dc = self._q(t["r_x"], X[m]), fw["x_reconstr_means"][m], fw["dec"][m]
            if self.binary[m]:
                da = (w / Bg) * (-x / (CE_EPS + xh) + (1 - x) / (CE_EPS + 1 - xh)) * xh * (1 - xh)
            else:
                da = w * (xh - x)                                   # l2_loss is a batch SUM: no 1/B
            if self.network_architectures[m]["hidden_conv"]:
                G[15] = dc["g2"].T @ da; G[16] = da.sum(0)
                do4 = (da @ P[15].T).reshape(dc["o4"].shape)
                d = do4 * dc["o4"] * (1 - dc["o4"])
                G[14] = d.sum((0, 1, 2)); d, G[13] = conv2d_transpose_bwd(dc["o3"], P[13], 2, "SAME", d)
                d = d * dc["o3"] * (1 - dc["o3"])
                G[12] = d.sum((0, 1, 2)); d, G[11] = conv2d_transpose_bwd(dc["o2"], P[11], 2, "SAME", d)
                d = d * dc["o2"] * (1 - dc["o2"])
                G[10] = d.sum((0, 1, 2)); d, G[9] = conv2d_transpose_bwd(dc["o1"], P[9], 1, "VALID", d)
                d = d * dc["o1"] * (1 - dc["o1"])
                G[8] = d.sum((0, 1, 2)); d, G[7] = conv2d_transpose_bwd(dc["z2"], P[7], 1, "VALID", d)
                dz = d.reshape(B, self.n_z)
            else:
                da = self._q(t["r_da"], da)
                G[12] = dc["g2"].T @ da; G[13] = da.sum(0)
                d = self._q(t["r_dg2"], (da @ self._q(t["d_o"], P[12]).T) * ag(m, "g2", dc["g2"]))
                G[10] = dc["g1"].T @ d; G[11] = d.sum(0)
                d = self._q(t["r_dg1"], (d @ self._q(t["d_d2"], P[10]).T) * ag(m, "g1", dc["g1"]))
                G[8] = dc["z"].T @ d; G[9] = d.sum(0)
                dz = d @ self._q(t["d_d1"], P[8]).T
            dz_list.append(dz)
        # latent: reparameterisation + prior KL + association KL (SURVEY 3.2 analytic gradients)
        dmu = [None] * M
        dlv = [None] * M
        for m in range(M):
            mu, lv, w = fw["z_means"][m], fw["z_log_sigma_sqs"][m], self.weights[m]
            s = np.sqrt(np.exp(lv))
            dmu[m] = dz_list[m] + (w / Bg) * mu
            dlv[m] = dz_list[m] * eps * 0.5 * s + (w / Bg) * 0.5 * (np.exp(lv) - 1.0)
        for i, j in itertools.combinations(range(M), 2):
            mi, mj = fw["z_means"][i], fw["z_means"][j]
            li, lj = fw["z_log_sigma_sqs"][i], fw["z_log_sigma_sqs"][j]
            dmu[i] = dmu[i] + lam * (mi - mj) * (np.exp(-li) + np.exp(-lj))
            dmu[j] = dmu[j] + lam * (mj - mi) * (np.exp(-li) + np.exp(-lj))
            dlv[i] = dlv[i] + lam * 0.5 * (np.exp(li - lj) - np.exp(lj - li) - (mi - mj) ** 2 * np.exp(-li))
            dlv[j] = dlv[j] + lam * 0.5 * (np.exp(lj - li) - np.exp(li - lj) - (mi - mj) ** 2 * np.exp(-lj))
        # encoders
        for m in range(M):
            P, G, ec = self.params[m], grads[m], fw["enc"][m]
            if self.network_architectures[m]["hidden_conv"]:
                G[3] = ec["h2"].T @ dmu[m]; G[4] = dmu[m].sum(0)
                G[5] = ec["h2"].T @ dlv[m]; G[6] = dlv[m].sum(0)
                d = (dmu[m] @ P[3].T + dlv[m] @ P[5].T).reshape(ec["l2shape"])
                d, G[2] = conv2d_bwd(ec["l1"], P[2], 1, "VALID", d)
                d, G[1] = conv2d_bwd(ec["l05"], P[1], 2, "SAME", d)
                _, G[0] = conv2d_bwd(ec["x2"], P[0], 2, "SAME", d)
            else:
                t = self._plan(m, B)
                dm_, dl_ = self._q(t["r_dhd"], dmu[m]), self._q(t["r_dhd"], dlv[m])
                G[4] = ec["h2"].T @ dm_; G[5] = dm_.sum(0)
                G[6] = ec["h2"].T @ dl_; G[7] = dl_.sum(0)
                d = (dm_ @ self._q(t["d_hd"], P[4]).T + dl_ @ self._q(t["d_hd"], P[6]).T) * ag(m, "h2", ec["h2"])
                d = self._q(t["r_dh2"], d)
                G[2] = ec["h1"].T @ d; G[3] = d.sum(0)
                d = self._q(t["r_dh1"], (d @ self._q(t["d_e2"], P[2]).T) * ag(m, "h1", ec["h1"]))
                G[0] = ec["x"].T @ d; G[1] = d.sum(0)
        probes = dict(fw)
        probes.update(ls)
        probes["d_z_means"] = dmu
        probes["d_z_log_sigma_sqs"] = dlv
        return ls["cost"], grads, probes

    # ---- Adam (TensorFlow ApplyAdam, third-party: lr_t form, epsilon outside the bias correction) ----
    def adam_step(self, grads):
        self.t += 1
        lr_t = self.learning_rate * np.sqrt(1.0 - ADAM_BETA2 ** self.t) / (1.0 - ADAM_BETA1 ** self.t)
        for ps, ms, vs, gs in zip(self.params, self.m, self.v, grads):
            for p, m, v, g in zip(ps, ms, vs, gs):
                m *= ADAM_BETA1; m += (1 - ADAM_BETA1) * g
                v *= ADAM_BETA2; v += (1 - ADAM_BETA2) * g * g
                p -= lr_t * m / (np.sqrt(v) + ADAM_EPS)

    # ---- reference API (vae_assoc.py:378-425) -------------------------------------------
    def _eps(self, eps):
        return self.rng.normal(size=(self.batch_size, self.n_z)) if eps is None else eps

    def partial_fit(self, X, eps=None):                                                          # :378-386
        cost, grads, _ = self.loss_and_grads(X, self._eps(eps))
        self.adam_step(grads)
        return cost

    def evaluate_cost(self, X, eps=None):                                                        # :388-391
        X = [np.asarray(x, dtype=self.dtype) for x in X]
        return self.loss(X, self.forward(X, self._eps(eps)))["cost"]

    def transform(self, X, sens_idx=None):                                                       # :393-403
        if sens_idx is None:
            return [self.encode(m, np.asarray(x, dtype=self.dtype))[0] for m, x in enumerate(X)]
        assert sens_idx < len(self.network_architectures)
        return self.encode(sens_idx, np.asarray(X, dtype=self.dtype))[0]

    def generate(self, z_mu=None):                                                               # :405-419
        if z_mu is None:
            z_mu = self.rng.normal(size=(self.batch_size, self.n_z))
        z_mu = np.asarray(z_mu, dtype=self.dtype)
        return [self.decode(m, z_mu)[0] for m in range(len(self.network_architectures))]

    def reconstruct(self, X, eps=None):                                                          # :421-425
        out = []
        for m, x in enumerate(X):
            mu, lv, _ = self.encode(m, np.asarray(x, dtype=self.dtype))
            e = self._eps(None if eps is None else eps[m] if isinstance(eps, (list, tuple)) else eps)
            out.append(self.decode(m, mu + np.sqrt(np.exp(lv)) * e)[0])
        return out


def reference_archs(n_z=4, conv=False, scaled=False):
    """Architecture dicts of vae_assoc_ujichar_img_jnt.py:53-71 (dense) / :72-80 (conv, commented out there)."""
    if scaled:
        img = dict(scope="image", hidden_conv=False, n_hidden_recog_1=2048, n_hidden_recog_2=2048,
                   n_hidden_gener_1=2048, n_hidden_gener_2=2048, n_input=784, n_z=n_z)
        jnt = dict(scope="joint", hidden_conv=False, n_hidden_recog_1=2048, n_hidden_recog_2=2048,
                   n_hidden_gener_1=2048, n_hidden_gener_2=2048, n_input=147, n_z=n_z)
        return [img, jnt]
    if conv:
        img = dict(scope="image", hidden_conv=True, n_hidden_recog_1=16, n_hidden_recog_2=64,
                   n_hidden_gener_1=64, n_hidden_gener_2=16, n_input=28 * 28, n_z=n_z)
    else:
        img = dict(scope="image", hidden_conv=False, n_hidden_recog_1=500, n_hidden_recog_2=500,
                   n_hidden_gener_1=500, n_hidden_gener_2=500, n_input=784, n_z=n_z)
    jnt = dict(scope="joint", hidden_conv=False, n_hidden_recog_1=200, n_hidden_recog_2=200,
               n_hidden_gener_1=200, n_hidden_gener_2=200, n_input=147, n_z=n_z)
    return [img, jnt]
