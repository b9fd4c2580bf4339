"""Host-side mirror of the reference's model surface (/root/reference/vae_assoc.py) over libvaeassoc.

Same names, argument meaning and error behaviour as the reference class (`AssocVariationalAutoEncoder`,
vae_assoc.py:20-463) and trainer (`train`, vae_assoc.py:498-583), so vae_assoc_ujichar_img_jnt.py and
vae_assoc_model_viewer.py stay valid callers.  Every arithmetic op goes through the C-ABI
(include/vaeassoc.h) into hand-written sm_100a kernels; PyTorch is only the device-memory / stream /
torch.distributed shell.  There is no CPU fallback: construction raises without the built library or a B200.
"""
import ctypes as C
import datetime
import os
import time

import numpy as np

from . import _lib as L


def _fct_name(transfer_fct):
    """The reference passes tf.nn.relu / tf.nn.softplus callables (vae_assoc.py:26,502); accept those (by
    __name__), the shim's functions, or plain strings."""
    name = transfer_fct if isinstance(transfer_fct, str) else getattr(transfer_fct, "__name__", "")
    name = name.lower()
    if name not in ("relu", "softplus"):
        raise ValueError("transfer_fct must be relu or softplus, got %r" % (transfer_fct,))
    return name


def softplus(x):   # stands in for the reference's default `tf.nn.softplus` (vae_assoc.py:26); a token, not a kernel
    raise RuntimeError("softplus is a selector token; the activation runs inside the CUDA kernels")


def relu(x):
    raise RuntimeError("relu is a selector token; the activation runs inside the CUDA kernels")


def xavier_init(fan_in, fan_out, constant=1, rng=None):
    """vae_assoc.py:11-18 -- U(+-c*sqrt(6/(fan_in+fan_out))), [fan_in, fan_out] fp32 (host-side, seeded)."""
    rng = np.random if rng is None else rng
    high = constant * np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-high, high, size=(fan_in, fan_out)).astype(np.float32)


class VaeAssocError(RuntimeError):
    pass


class AssocVariationalAutoEncoder(object):
    """Drop-in for the reference class (vae_assoc.py:20).  Extra keyword-only arguments select the precision
    (`"tf32"`: tcgen05 tensor cores, `"fp32"`: SIMT FFMA), the device and the RNG seeds."""

    def __init__(self, network_architectures, binary=True, transfer_fct=softplus, weights=1.0, assoc_lambda=1.0,
                 learning_rate=0.001, batch_size=100, precision="tf32", device=None, seed=0, eps_seed=0,
                 use_graph=True, global_batch=None, global_row0=0):
        import torch
        self._torch = torch
        self.network_architectures = network_architectures
        self.assoc_lambda = assoc_lambda
        n = len(network_architectures)
        if type(binary) is list:                                   # vae_assoc.py:31-35
            assert len(binary) == n
            self.binary = binary
        else:
            self.binary = [binary] * n
        if type(weights) is list:                                  # vae_assoc.py:37-41
            assert len(weights) == n
            self.weights = weights
        else:
            self.weights = [weights] * n
        self.transfer_fct = transfer_fct
        self.learning_rate = learning_rate
        self.batch_size = batch_size
        self.n_z = network_architectures[0]["n_z"]                 # vae_assoc.py:89
        self.precision = precision
        if n > L.MAX_MODALITIES:
            raise ValueError("at most %d modalities" % L.MAX_MODALITIES)
        if not torch.cuda.is_available():
            raise VaeAssocError("no CUDA device: the associated-VAE train step has no CPU fallback")
        self._lib = L.load()
        self._device = torch.cuda.current_device() if device is None else int(device)
        cfg = L.Config()
        cfg.abi_version = L.ABI_VERSION
        cfg.n_modalities = n
        cfg.batch_size = int(batch_size)
        cfg.n_z = int(self.n_z)
        cfg.transfer_fct = L.SOFTPLUS if _fct_name(transfer_fct) == "softplus" else L.RELU
        cfg.precision = {"fp32": L.FP32, "tf32": L.TF32}[precision]
        cfg.device = self._device
        cfg.use_graph = 1 if use_graph else 0
        cfg.assoc_lambda = float(assoc_lambda)
        cfg.learning_rate = float(learning_rate)
        cfg.beta1, cfg.beta2, cfg.adam_epsilon = 0.9, 0.999, 1e-8   # tf.train.AdamOptimizer defaults (:373-374)
        cfg.global_batch = int(global_batch or 0)
        cfg.global_row0 = int(global_row0)
        cfg.eps_seed = int(eps_seed) & 0xFFFFFFFF
        for m, na in enumerate(network_architectures):
            mod = cfg.mod[m]
            mod.n_input = int(na["n_input"])
            mod.n_hidden_recog_1 = int(na["n_hidden_recog_1"]); mod.n_hidden_recog_2 = int(na["n_hidden_recog_2"])
            mod.n_hidden_gener_1 = int(na["n_hidden_gener_1"]); mod.n_hidden_gener_2 = int(na["n_hidden_gener_2"])
            mod.hidden_conv = 1 if na.get("hidden_conv", False) else 0
            mod.binary = 1 if self.binary[m] else 0
            mod.weight = float(self.weights[m])
            mod.scope = str(na.get("scope", "")).encode()[:31]      # tf.variable_scope(scope), vae_assoc.py:168,248
            assert na["n_z"] == self.n_z, "all modalities share one latent size (vae_assoc.py:89-91)"
        self._cfg = cfg
        h = L.Handle()
        if self._lib.vaeassoc_create(C.byref(cfg), C.byref(h)) != 0:
            raise VaeAssocError(self._lib.vaeassoc_last_error(None).decode())
        self._h = h
        self._dev = torch.device("cuda", self._device)
        self._bind_stream()
        self._tensors = []
        for i in range(self._lib.vaeassoc_num_tensors(h)):
            ti = L.TensorInfo()
            self._check(self._lib.vaeassoc_layout_query(h, i, C.byref(ti)))
            self._tensors.append(ti)
        self._rng = np.random.RandomState(seed)
        self._init_weights()
        self._prior_draws = 0
        self._keep = [None, None]       # host arrays of the two most recent pipelined submits (kept alive across the DMA)
        self._world = 1
        from . import tf_shim
        tf_shim.register(self)          # tf.all_variables() / tf.reset_default_graph() see this model (weak reference)

    # ---- plumbing --------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise VaeAssocError(self._lib.vaeassoc_last_error(self._h).decode())

    def _bind_stream(self):
        # torch's default stream has handle 0, which the C-ABI reads as "the handle's own stream": name the legacy
        # default stream explicitly (cudaStreamLegacy == 0x1) so that torch copies/events order with our kernels
        s = self._torch.cuda.current_stream(self._dev).cuda_stream or 1
        if s != getattr(self, "_bound_stream", None):
            self._check(self._lib.vaeassoc_set_stream(self._h, C.c_void_p(s)))
            self._bound_stream = s

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_peer", False):
                try:
                    self.detach_peers()
                except Exception:
                    pass
            self._lib.vaeassoc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _init_weights(self):
        """Reference initialisers (vae_assoc.py:185-215,257-300): Xavier-uniform weights, zero biases."""
        for i, ti in enumerate(self._tensors):
            shape = tuple(ti.shape[:ti.ndim])
            role = ti.role.decode()
            if ti.ndim == 1:
                w = np.zeros(shape, np.float32)                      # every bias starts at zero (deconv.py:113 too)
            elif role in ("C1", "C2", "C3"):
                # conv_2d -> weight_variable: tf.truncated_normal(stddev=0.1) (vae_assoc.py:471-473,482)
                w = self._rng.normal(size=shape)
                bad = np.abs(w) > 2.0
                while bad.any():
                    w[bad] = self._rng.normal(size=int(bad.sum()))
                    bad = np.abs(w) > 2.0
                w = (0.1 * w).astype(np.float32)
            elif role in ("D1", "D2", "D3", "D4"):
                # deconv2d: prettytensor xavier_init(out*k*k, in*k*k), filter [k,k,out,in] (deconv.py:78-84)
                k2 = shape[0] * shape[1]
                lim = np.sqrt(6.0 / (shape[2] * k2 + shape[3] * k2))
                w = self._rng.uniform(-lim, lim, size=shape).astype(np.float32)
            else:
                w = xavier_init(shape[0], shape[1], rng=self._rng)
            self._set(L.PARAMS, i, w)

    def _set(self, which, i, arr):
        ti = self._tensors[i]
        a = np.ascontiguousarray(arr, dtype=np.float32).reshape(int(ti.rows), int(ti.cols))
        self._check(self._lib.vaeassoc_tensor_set(self._h, which, i, a.ctypes.data_as(C.c_void_p)))

    def _get(self, which, i):
        ti = self._tensors[i]
        a = np.empty((int(ti.rows), int(ti.cols)), np.float32)
        self._check(self._lib.vaeassoc_tensor_get(self._h, which, i, a.ctypes.data_as(C.c_void_p)))
        return a.reshape(tuple(ti.shape[:ti.ndim]))

    # ---- parameter / optimiser-state access (tf.Variable + tf.train.Saver surface) ---------------------
    def variable_names(self):
        return [t.name.decode() for t in self._tensors]

    def variable_roles(self):
        return [(t.modality, t.role.decode()) for t in self._tensors]

    def get_params(self):
        return [self._get(L.PARAMS, i) for i in range(len(self._tensors))]

    def set_params(self, params):
        flat = [p for ps in params for p in ps] if isinstance(params[0], (list, tuple)) else list(params)
        assert len(flat) == len(self._tensors)
        for i, p in enumerate(flat):
            self._set(L.PARAMS, i, p)

    def get_grads(self):
        return [self._get(L.GRADS, i) for i in range(len(self._tensors))]

    def get_adam_state(self):
        step = C.c_int64()
        self._check(self._lib.vaeassoc_step_get(self._h, C.byref(step)))
        return ([self._get(L.ADAM_M, i) for i in range(len(self._tensors))],
                [self._get(L.ADAM_V, i) for i in range(len(self._tensors))], int(step.value))

    def set_adam_state(self, m, v, step):
        for i in range(len(self._tensors)):
            self._set(L.ADAM_M, i, m[i])
            self._set(L.ADAM_V, i, v[i])
        self._check(self._lib.vaeassoc_step_set(self._h, int(step)))

    def set_precision(self, precision):
        self._check(self._lib.vaeassoc_set_precision(self._h, {"fp32": L.FP32, "tf32": L.TF32}[precision]))
        self.precision = precision

    def guard_check(self):
        """(number of guard regions, number corrupted): every device buffer of the handle ends in a 256-byte pattern that
        no kernel may write (the library's own bounds check; compute-sanitizer is closed on the GPU pool)."""
        n, bad = C.c_int64(), C.c_int64()
        self._check(self._lib.vaeassoc_debug_guard_check(self._h, C.byref(n), C.byref(bad)))
        return int(n.value), int(bad.value)

    def launch_count(self):
        return int(self._lib.vaeassoc_launch_count(self._h))

    # ---- argument marshalling -----------------------------------------------------------------------
    def _is_device(self, X):
        return all(self._torch.is_tensor(x) and x.is_cuda for x in X)

    def _dev_args(self, X):
        ptrs = (C.c_void_p * L.MAX_MODALITIES)()
        lds = (C.c_int64 * L.MAX_MODALITIES)()
        keep = []
        for m, x in enumerate(X):
            na = self.network_architectures[m]
            if x.dtype != self._torch.float32 or x.stride(1) != 1:
                x = x.float().contiguous()
            assert tuple(x.shape) == (self.batch_size, na["n_input"]), \
                "modality %d: expected %s, got %s (the batch size is static, vae_assoc.py:90)" % (
                    m, (self.batch_size, na["n_input"]), tuple(x.shape))
            keep.append(x)
            ptrs[m] = x.data_ptr()
            lds[m] = x.stride(0)
        return ptrs, lds, keep

    def _host_args(self, X):
        ptrs = (C.c_void_p * L.MAX_MODALITIES)()
        keep = []
        for m, x in enumerate(X):
            na = self.network_architectures[m]
            if self._torch.is_tensor(x):
                x = x.detach().cpu().numpy()
            a = np.ascontiguousarray(x, dtype=np.float32)
            assert a.shape == (self.batch_size, na["n_input"]), \
                "modality %d: expected %s, got %s (the batch size is static, vae_assoc.py:90)" % (
                    m, (self.batch_size, na["n_input"]), a.shape)
            keep.append(a)
            ptrs[m] = a.ctypes.data
        return ptrs, keep

    def _eps_dev(self, eps):
        if eps is None:
            return None, None
        t = self._torch
        e = eps if t.is_tensor(eps) else t.as_tensor(np.asarray(eps, dtype=np.float32))
        e = e.to(self._dev, t.float32).contiguous()
        assert tuple(e.shape) == (self.batch_size, self.n_z)
        return C.c_void_p(e.data_ptr()), e

    def _to_dev(self, X):
        t = self._torch
        return [x if (t.is_tensor(x) and x.is_cuda) else t.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(self._dev)
                for x in X]

    # ---- the reference API -------------------------------------------------------------------------
    def partial_fit(self, X, eps=None):
        """Train model based on mini-batch of input data.  Return cost of mini-batch.  (vae_assoc.py:378-386)

        `X` = list of per-modality [batch_size, n_input] arrays (numpy -> host path with H2D inside the call,
        CUDA tensors -> device path).  `eps` optionally injects the reparameterisation noise (parity tests)."""
        self._bind_stream()
        cost = C.c_float()
        if self._is_device(X):
            ptrs, lds, keep = self._dev_args(X)
            ep, ekeep = self._eps_dev(eps)
            self._check(self._lib.vaeassoc_train_step(self._h, ptrs, lds, ep))
            self._check(self._lib.vaeassoc_cost_read(self._h, C.byref(cost)))
        else:
            ptrs, keep = self._host_args(X)
            e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float32)
            self._check(self._lib.vaeassoc_partial_fit_host(
                self._h, ptrs, None if e is None else e.ctypes.data_as(C.c_void_p), C.byref(cost)))
        return np.float32(cost.value)

    def partial_fit_async(self, X, eps=None):
        """Same step without the per-step host synchronisation the reference pays (vae_assoc.py:383-386);
        the cost lands in the device-side history (see `cost_history`)."""
        self._bind_stream()
        if self._is_device(X):
            ptrs, lds, keep = self._dev_args(X)
            ep, ekeep = self._eps_dev(eps)
            self._check(self._lib.vaeassoc_train_step(self._h, ptrs, lds, ep))
        else:
            ptrs, keep = self._host_args(X)
            e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float32)
            self._check(self._lib.vaeassoc_submit_host(self._h, ptrs, None if e is None else e.ctypes.data_as(C.c_void_p)))
            # the H2D copies of PINNED arrays are asynchronous DMA: keep the arrays of the last two submits alive (the
            # library double-buffers, so older copies have completed); callers that REFILL a pinned buffer must call
            # wait_uploaded(k) first (include/vaeassoc.h, vaeassoc_submit_host)
            k = int(self._lib.vaeassoc_submit_count(self._h)) - 1
            self._keep[k & 1] = (keep, e)
            return k

    def compute_gradients(self, X, eps=None):
        """Forward + backward without the Adam update; returns the cost.  Gradients: `get_grads()`."""
        self._bind_stream()
        X = self._to_dev(X)
        ptrs, lds, keep = self._dev_args(X)
        ep, ekeep = self._eps_dev(eps)
        self._check(self._lib.vaeassoc_grad_step(self._h, ptrs, lds, ep))
        cost = C.c_float()
        self._check(self._lib.vaeassoc_cost_read(self._h, C.byref(cost)))
        return np.float32(cost.value)

    def last_cost(self):
        cost = C.c_float()
        self._check(self._lib.vaeassoc_cost_read(self._h, C.byref(cost)))
        return np.float32(cost.value)

    def cost_history(self, first_step, n):
        out = np.empty(int(n), np.float32)
        self._check(self._lib.vaeassoc_cost_history(self._h, int(first_step), int(n), out.ctypes.data_as(C.c_void_p)))
        return out

    def submit_costs(self, first_submit, n):
        """Costs of host-path submits [first, first+n), read from the pinned ring the steps D2H into."""
        out = np.empty(int(n), np.float32)
        self._check(self._lib.vaeassoc_submit_costs(self._h, int(first_submit), int(n), out.ctypes.data_as(C.c_void_p)))
        return out

    def upload_dataset(self, rows):
        """Uploads a whole [N, sum n_input] training matrix once (row pitch padded to 16 bytes); returns the device
        tensor to pass to `partial_fit_indexed`.  The reference keeps the data set in host RAM and copies every batch
        through feed_dict (vae_assoc.py:541-543,574); here only the batch's row indices cross PCIe."""
        t = self._torch
        rows = np.asarray(rows)
        width = sum(int(na["n_input"]) for na in self.network_architectures)
        assert rows.ndim == 2 and rows.shape[1] == width, (rows.shape, width)
        dev = t.zeros((rows.shape[0], (width + 3) // 4 * 4), dtype=t.float32, device=self._dev)
        dev[:, :width].copy_(t.as_tensor(np.ascontiguousarray(rows, dtype=np.float32)))
        return dev

    def partial_fit_indexed(self, data_dev, indices, eps=None):
        """One pipelined train step on rows `indices` (host int64 [batch_size]) of a device-resident data set."""
        self._bind_stream()
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        assert idx.shape == (self.batch_size,)
        e = None if eps is None else np.ascontiguousarray(eps, dtype=np.float32)
        self._check(self._lib.vaeassoc_submit_indexed(self._h, C.c_void_p(data_dev.data_ptr()), data_dev.stride(0),
                                                      data_dev.shape[0], idx.ctypes.data_as(C.c_void_p),
                                                      None if e is None else e.ctypes.data_as(C.c_void_p)))
        return int(self._lib.vaeassoc_submit_count(self._h)) - 1

    def wait_uploaded(self, submit_index):
        """Blocks until the host->device copies of pipelined submit `submit_index` (the value partial_fit_async returned)
        have completed; after that its (pinned) host buffers may be refilled."""
        self._check(self._lib.vaeassoc_upload_wait(self._h, int(submit_index)))

    def synchronize(self):
        self._check(self._lib.vaeassoc_stream_sync(self._h))

    def relu_masks(self, m):
        """The relu sign masks the tensor-core path stored for modality m at the most recent step, as boolean arrays
        {"h1","h2","g1","g2"} of shape [batch_size, width] (TF ReluGrad's `y > 0`).  Parity tests feed them to the oracle."""
        na = self.network_architectures[m]
        out = {}
        for layer, (name, width) in enumerate([("h1", na["n_hidden_recog_1"]), ("h2", na["n_hidden_recog_2"]),
                                               ("g1", na["n_hidden_recog_1"]), ("g2", na["n_hidden_recog_2"])]):
            wpr = (width + 31) // 32
            buf = np.empty(self.batch_size * wpr, np.uint32)
            n, w = C.c_int64(), C.c_int64()
            self._check(self._lib.vaeassoc_probe_mask(self._h, layer, m, buf.ctypes.data_as(C.c_void_p), buf.size,
                                                      C.byref(n), C.byref(w)))
            words = buf.reshape(self.batch_size, wpr)
            bits = (words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1
            out[name] = bits.reshape(self.batch_size, wpr * 32)[:, :width].astype(bool)
        return out

    def evaluate_cost(self, X, eps=None):
        """vae_assoc.py:388-391"""
        self._bind_stream()
        X = self._to_dev(X)
        ptrs, lds, keep = self._dev_args(X)
        ep, ekeep = self._eps_dev(eps)
        cost = C.c_float()
        self._check(self._lib.vaeassoc_eval_cost(self._h, ptrs, lds, ep, C.byref(cost)))
        return np.float32(cost.value)

    def _encode(self, m, x):
        t = self._torch
        x = self._to_dev([x])[0]
        if x.dtype != t.float32 or x.stride(1) != 1:
            x = x.float().contiguous()
        assert tuple(x.shape) == (self.batch_size, self.network_architectures[m]["n_input"])
        mu = t.empty((self.batch_size, self.n_z), dtype=t.float32, device=self._dev)
        self._check(self._lib.vaeassoc_encode(self._h, m, C.c_void_p(x.data_ptr()), x.stride(0),
                                              C.c_void_p(mu.data_ptr()), None))
        return mu.cpu().numpy()

    def _infer_host(self, kind, modality, X, z_or_eps, widths):
        """transform / generate / reconstruct for host arrays: one graph launch + one D2H (vaeassoc_infer_host)."""
        M = len(self.network_architectures)
        ptrs = (C.c_void_p * L.MAX_MODALITIES)()
        outs = (C.c_void_p * L.MAX_MODALITIES)()
        keep, res = [], [None] * M
        mods = range(M) if modality < 0 else [modality]
        for m in mods:
            if X is not None:
                a = np.ascontiguousarray(X[m], dtype=np.float32)
                assert a.shape == (self.batch_size, self.network_architectures[m]["n_input"]), \
                    "modality %d: expected %s, got %s (the batch size is static, vae_assoc.py:90)" % (
                        m, (self.batch_size, self.network_architectures[m]["n_input"]), a.shape)
                keep.append(a)
                ptrs[m] = a.ctypes.data
            res[m] = np.empty((self.batch_size, widths[m]), np.float32)
            outs[m] = res[m].ctypes.data
        zp = None
        if z_or_eps is not None:
            z = np.ascontiguousarray(z_or_eps, dtype=np.float32)
            assert z.shape == (self.batch_size, self.n_z), \
                "z / eps must be [batch_size, n_z] (the reference's callers pad to a full batch, baxter_vae_assoc_writer.py:142-145)"
            keep.append(z)
            zp = z.ctypes.data_as(C.c_void_p)
        self._check(self._lib.vaeassoc_infer_host(self._h, kind, modality, ptrs if X is not None else None, zp, outs))
        return res

    def _all_host(self, X):
        t = self._torch
        return not any(t.is_tensor(x) and x.is_cuda for x in X)

    def transform(self, X, sens_idx=None):
        """Transform data by mapping it into the latent space (z_mean).  (vae_assoc.py:393-403)"""
        self._bind_stream()
        M = len(self.network_architectures)
        if sens_idx is None:
            if self._all_host(X):
                return self._infer_host(0, -1, [x.cpu().numpy() if self._torch.is_tensor(x) else x for x in X], None,
                                        [self.n_z] * M)
            return [self._encode(m, x) for m, x in enumerate(X)]
        assert sens_idx < M
        if self._all_host([X]):
            Xs = [None] * M
            Xs[sens_idx] = X.cpu().numpy() if self._torch.is_tensor(X) else X
            return self._infer_host(0, sens_idx, Xs, None, [self.n_z] * M)[sens_idx]
        return self._encode(sens_idx, X)

    def generate(self, z_mu=None):
        """Generate data by sampling from latent space; z feeds the decoders directly.  (vae_assoc.py:405-419)"""
        self._bind_stream()
        t = self._torch
        widths = [na["n_input"] for na in self.network_architectures]
        if z_mu is not None and not (t.is_tensor(z_mu) and z_mu.is_cuda):
            return self._infer_host(1, -1, None, z_mu.cpu().numpy() if t.is_tensor(z_mu) else z_mu, widths)
        if z_mu is None:
            # the reference draws np.random.normal((batch_size, n_z)) (:414); here Philox on the device
            z = t.empty((self.batch_size, self.n_z), dtype=t.float32, device=self._dev)
            self._check(self._lib.vaeassoc_philox_normal(self._h, self._cfg.eps_seed, 5, 0, self.batch_size, self.n_z,
                                                         self._prior_draws, C.c_void_p(z.data_ptr())))
            self._prior_draws += 1
        else:
            z = z_mu.to(self._dev, t.float32).contiguous()
        assert tuple(z.shape) == (self.batch_size, self.n_z), \
            "z_mu must be [batch_size, n_z] (the reference's callers pad to a full batch, baxter_vae_assoc_writer.py:142-145)"
        out = []
        for m, na in enumerate(self.network_architectures):
            xh = t.empty((self.batch_size, na["n_input"]), dtype=t.float32, device=self._dev)
            self._check(self._lib.vaeassoc_decode(self._h, m, C.c_void_p(z.data_ptr()), C.c_void_p(xh.data_ptr())))
            out.append(xh.cpu().numpy())
        return out

    def reconstruct(self, X, eps=None):
        """Use VAE to reconstruct given data: one encode+sample+decode per modality.  (vae_assoc.py:421-425)"""
        self._bind_stream()
        t = self._torch
        shared_eps = eps is None or not isinstance(eps, (list, tuple))
        if self._all_host(X) and shared_eps and not (t.is_tensor(eps) and eps.is_cuda):
            e = eps.cpu().numpy() if t.is_tensor(eps) else eps
            return self._infer_host(2, -1, [x.cpu().numpy() if t.is_tensor(x) else x for x in X], e,
                                    [na["n_input"] for na in self.network_architectures])
        out = []
        for m, x in enumerate(X):
            na = self.network_architectures[m]
            x = self._to_dev([x])[0]
            if x.dtype != t.float32 or x.stride(1) != 1:
                x = x.float().contiguous()
            assert tuple(x.shape) == (self.batch_size, na["n_input"])
            e = None if eps is None else (eps[m] if isinstance(eps, (list, tuple)) else eps)
            ep, ekeep = self._eps_dev(e)
            xh = t.empty((self.batch_size, na["n_input"]), dtype=t.float32, device=self._dev)
            self._check(self._lib.vaeassoc_reconstruct(self._h, m, C.c_void_p(x.data_ptr()), x.stride(0), ep,
                                                       C.c_void_p(xh.data_ptr())))
            out.append(xh.cpu().numpy())
        return out

    # ---- probe points (the author's commented-out debugging fetches, vae_assoc.py:545-571) ----------------
    def _probe(self, kind, m, n):
        buf = np.empty(int(n), np.float32)
        wrote = C.c_int64()
        self._check(self._lib.vaeassoc_probe_get(self._h, kind, m, buf.ctypes.data_as(C.c_void_p), int(n), C.byref(wrote)))
        return buf[:wrote.value]

    def _probe_list(self, kind, cols=None):
        out = []
        for m, na in enumerate(self.network_architectures):
            c = self.n_z if cols is None else cols(na)
            out.append(self._probe(kind, m, self.batch_size * c).reshape(self.batch_size, c))
        return out

    @property
    def z_means(self):
        return self._probe_list(L.PROBE_Z_MEAN)

    @property
    def z_log_sigma_sqs(self):
        return self._probe_list(L.PROBE_Z_LOG_SIGMA_SQ)

    @property
    def z_array(self):
        return self._probe_list(L.PROBE_Z)

    @property
    def x_reconstr_means(self):
        return self._probe_list(L.PROBE_X_RECONSTR_MEAN, cols=lambda na: na["n_input"])

    @property
    def d_z_means(self):
        return self._probe_list(L.PROBE_D_Z_MEAN)

    @property
    def d_z_log_sigma_sqs(self):
        return self._probe_list(L.PROBE_D_Z_LOG_SIGMA_SQ)

    @property
    def vae_reconstr_losses(self):
        out = []
        for m in range(len(self.network_architectures)):
            r = self._probe(L.PROBE_RECONSTR_LOSS, m, self.batch_size)
            out.append(r if self.binary[m] else np.float32(r[0]))
        return out

    @property
    def vae_latent_losses(self):
        return [self._probe(L.PROBE_LATENT_LOSS, m, self.batch_size) for m in range(len(self.network_architectures))]

    @property
    def vae_costs(self):
        return [np.float32(self._probe(L.PROBE_VAE_COST, m, 1)[0]) for m in range(len(self.network_architectures))]

    @property
    def assoc_costs(self):
        if len(self.network_architectures) < 2:
            return []
        return [np.float32(self._probe(L.PROBE_ASSOC_COST, 0, 1)[0])]   # summed over modality pairs

    @property
    def cost(self):
        return self.last_cost()

    @property
    def last_eps(self):
        return self._probe(L.PROBE_EPS, 0, self.batch_size * self.n_z).reshape(self.batch_size, self.n_z)

    # ---- synthetic paired batches (replaces dataset.py / utils.py for benchmarks) ------------------------
    def synth_batch(self, row0, n_rows=None, data_seed=0, proj_seed=1):
        """Device-resident synthetic pairs for global rows [row0, row0+n_rows): list of CUDA tensors."""
        self._bind_stream()
        t = self._torch
        n_rows = self.batch_size if n_rows is None else int(n_rows)
        outs = [t.empty((n_rows, na["n_input"]), dtype=t.float32, device=self._dev) for na in self.network_architectures]
        ptrs = (C.c_void_p * L.MAX_MODALITIES)()
        for m, o in enumerate(outs):
            ptrs[m] = o.data_ptr()
        self._check(self._lib.vaeassoc_synth_batch(self._h, int(data_seed), int(proj_seed), int(row0), n_rows, ptrs))
        return outs

    def philox_normal(self, seed, tag, row0, n_rows, n_cols, step=0):
        t = self._torch
        out = t.empty((n_rows, n_cols), dtype=t.float32, device=self._dev)
        self._check(self._lib.vaeassoc_philox_normal(self._h, int(seed), int(tag), int(row0), int(n_rows), int(n_cols),
                                                     int(step), C.c_void_p(out.data_ptr())))
        return out

    # ---- data parallelism ----------------------------------------------------------------------------
    def init_data_parallel(self, peer=None):
        """Join the data-parallel job of the current torch.distributed process group (one process per GPU): an NCCL
        communicator for the set-up / evaluation collectives and, with `peer` (default: on, VAEASSOC_DP_PEER=0 turns it
        off), the NVLink peer-memory train step -- one kernel that reduce-scatters the gradients, runs Adam on the owned
        shard and all-gathers the parameters (csrc/peer_adam.cu) -- in place of ncclAllReduce + replicated Adam."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        path = L.nccl_library_path().encode()
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            if self._lib.vaeassoc_comm_unique_id(path, buf) != 0:
                raise VaeAssocError(self._lib.vaeassoc_last_error(None).decode())
        ids = [bytes(buf)]
        dist.broadcast_object_list(ids, src=0)
        buf = (C.c_ubyte * 128).from_buffer_copy(ids[0])
        self._check(self._lib.vaeassoc_comm_init(self._h, path, buf, rank, world))
        self._world = world
        # replicas must start identical (the all-reduce only averages gradients): rank 0's parameters, Adam slots and step
        self.sync_replicas()
        if peer is None:
            peer = os.environ.get("VAEASSOC_DP_PEER", "1") != "0"
        self._peer = False
        if peer and 2 <= world <= 8:
            self._attach_peers(dist, world)

    def _attach_peers(self, dist, world):
        """Peer-memory data-parallel step (csrc/peer_adam.cu).  First choice: symmetric memory from
        torch.distributed._symmetric_memory (every rank's flat buffers mapped everywhere + an NVSwitch multicast address:
        the gradients are reduced INSIDE the switch, the parameters broadcast with one store); second: cudaIpc handles of
        the library's own allocation (plain peer loads / stores); else the NCCL all-reduce schedule.  All ranks end up in
        the same mode: each stage is agreed on by an all-gather of the per-rank outcome."""
        def agreed(ok):
            oks = [None] * world
            dist.all_gather_object(oks, bool(ok))      # also the barrier between attach and the first step
            return all(oks)

        self._peer_error = None
        # measured on B200 (profiles/r2_dp_modes.md): the NVLS form moves 1/world of the bytes but its multimem stores and
        # the system fence behind them take longer than the plain peer stores at this size (5.7 MB of gradients): 0.350
        # against 0.334 ms per step at 2 ranks, equal at 8 -- so the plain peer form is the default, NVLS is opt-in
        if os.environ.get("VAEASSOC_DP_SYMMETRIC", "0") != "0":
            ok = False
            try:
                import torch.distributed._symmetric_memory as symm_mem
                t = self._torch
                n = int(self._lib.vaeassoc_arena_floats(self._h))
                buf = symm_mem.empty(n, dtype=t.float32, device=self._dev)
                hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
                ptrs = (C.c_void_p * 8)(*[int(p) for p in hdl.buffer_ptrs])      # kMaxPeers slots, rank order
                mc = int(hdl.multicast_ptr or 0)
                t.cuda.synchronize()
                ok = self._lib.vaeassoc_peer_attach_symmetric(self._h, C.c_void_p(buf.data_ptr()), ptrs,
                                                              C.c_void_p(mc) if mc else None) == 0
                if not ok:
                    self._peer_error = self._lib.vaeassoc_last_error(self._h).decode()
                self._symm = (buf, hdl)        # the library's flat buffers now live here: keep it until close()
            except Exception as e:             # no symmetric-memory support in this torch / driver / topology
                self._peer_error = "symmetric memory: %r" % (e,)
            if agreed(ok):
                self._peer = True
                return
            if ok:
                self._check(self._lib.vaeassoc_peer_detach(self._h))
        blob = (C.c_ubyte * L.PEER_BLOB_BYTES)()
        ok = self._lib.vaeassoc_peer_export(self._h, blob) == 0
        blobs = [None] * world
        dist.all_gather_object(blobs, bytes(blob) if ok else None)
        if ok and all(b is not None for b in blobs):
            allb = (C.c_ubyte * (L.PEER_BLOB_BYTES * world)).from_buffer_copy(b"".join(blobs))
            ok = self._lib.vaeassoc_peer_attach(self._h, allb) == 0
        else:
            ok = False
        if not ok:
            self._peer_error = self._lib.vaeassoc_last_error(self._h).decode()
        if agreed(ok):
            self._peer = True
        elif ok:
            self._check(self._lib.vaeassoc_peer_detach(self._h))

    def detach_peers(self):
        """Back to the NCCL all-reduce schedule (collective: every rank calls it; no peer may still be stepping)."""
        if getattr(self, "_peer", False):
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.synchronize()
                dist.barrier()
            self._check(self._lib.vaeassoc_peer_detach(self._h))
            self._peer = False

    @property
    def dp_mode(self):
        if getattr(self, "_world", 1) < 2:
            return "single"
        if not self._lib.vaeassoc_peer_active(self._h):
            return "nccl"
        return "peer-nvls" if self._lib.vaeassoc_peer_multicast(self._h) else "peer"

    def sync_replicas(self):
        """Broadcast rank 0's parameters / Adam state to every rank (call on ALL ranks, e.g. after a rank-0 restore_model)."""
        self._check(self._lib.vaeassoc_comm_sync_state(self._h))

    def check_communicator(self):
        """ncclCommGetAsyncError; raises VaeAssocError (after aborting the communicator) on an asynchronous NCCL failure."""
        self._check(self._lib.vaeassoc_comm_check(self._h))

    # ---- checkpoints (tf.train.Saver surface, vae_assoc.py:70,427-463) -------------------------------------
    def save_model(self, fname=None):
        if fname is None:
            ts = time.time()
            ckpt_fname = 'vae_assoc_' + datetime.datetime.fromtimestamp(ts).strftime('%Y_%m_%d_%H_%M_%S') + \
                '_batchsize_{}.ckpt'.format(self.batch_size)
        else:
            ckpt_fname = fname
        print('Saving model to {}...'.format(ckpt_fname))
        from . import checkpoint
        checkpoint.save(self, ckpt_fname)
        return

    def restore_model(self, folder=None, fname=None):
        """Failures print and return, like the reference (vae_assoc.py:447-462)."""
        model_folder = 'output' if folder is None else folder
        if os.path.isdir(model_folder) and os.path.exists(model_folder):
            if fname is None:
                files = [f for f in os.listdir(model_folder) if f.endswith('.ckpt')]
                if not files:
                    print('No valid model file.')
                    return
                model_file = files[-1]
            else:
                model_file = fname
            path = os.path.join(model_folder, model_file)
            if os.path.exists(path):
                print('Loading {}...'.format(path))
                from . import checkpoint
                checkpoint.load(self, path)
            else:
                print('Invalid or non-exist model file.')
        else:
            print('Invalid or non-exist model folder.')
        return


def train(data_sets, network_architectures, binary=True, weights=1.0, assoc_lambda=1e-5, learning_rate=0.001,
          batch_size=100, training_epochs=10, display_step=5, early_stop=False, device_data=False, **model_kwargs):
    """vae_assoc.py:498-583.  Same loop and return value `(model, avg_cost_hist)`; differences that do not change
    results: batches are uploaded through the pipelined host path and the per-step cost is read back in chunks
    from a pinned ring instead of synchronising every step (vae_assoc.py:383-386).  `device_data=True` uploads
    `data_sets.train` ONCE and gathers every batch on the device from the row indices `next_batch` would use
    (same batches, same RNG consumption): 8 bytes per pair cross PCIe instead of 3 724."""
    vae_assoc = AssocVariationalAutoEncoder(network_architectures, binary, transfer_fct=relu, weights=weights,
                                            assoc_lambda=assoc_lambda, learning_rate=learning_rate,
                                            batch_size=batch_size, **model_kwargs)
    n_samples = data_sets.train._data.shape[0]
    sens_indices = np.concatenate([[0], np.cumsum([na["n_input"] for na in network_architectures])])
    n_mod = len(network_architectures)
    avg_cost_hist = []
    valid_cost = None
    # the reference slices the batch into per-modality column views (:543) and feed_dict copies them; here the slices are
    # written into a ring of PINNED per-modality staging buffers (true async DMA, overlapping the previous step's compute);
    # a slot is refilled only after vaeassoc_upload_wait says its previous upload has completed
    torch = vae_assoc._torch
    n_slots = 3
    staging = [[torch.empty((batch_size, int(na["n_input"])), dtype=torch.float32, pin_memory=True).numpy()
                for na in network_architectures] for _ in range(n_slots)]
    slot_submit = [None] * n_slots

    def segment(batch_xs, slot=None):
        if slot is None:
            return [np.ascontiguousarray(batch_xs[:, sens_indices[i]:sens_indices[i + 1]], dtype=np.float32)
                    for i in range(n_mod)]
        if slot_submit[slot] is not None:
            vae_assoc.wait_uploaded(slot_submit[slot])
        for i in range(n_mod):
            np.copyto(staging[slot][i], batch_xs[:, sens_indices[i]:sens_indices[i + 1]], casting="same_kind")
        return staging[slot]

    train_dev = vae_assoc.upload_dataset(data_sets.train._rows) if device_data else None
    submits = 0
    for epoch in range(training_epochs):
        avg_cost = 0.
        total_batch = int(n_samples / batch_size)
        if early_stop:
            if epoch % early_stop == 0:
                curr_valid_cost = 0
                n_valid_batches = int(data_sets.validation._data.shape[0] / batch_size)
                for i in range(n_valid_batches):
                    batch_xs, _ = data_sets.validation.next_batch(batch_size)
                    curr_valid_cost += vae_assoc.evaluate_cost(segment(batch_xs)) / n_valid_batches
                print("Validation cost=", "{:.9f}".format(curr_valid_cost))
                if valid_cost is not None:
                    if curr_valid_cost > valid_cost:
                        print('Validation error increases. Early stop at epoch {} to prevent overfitting...'.format(epoch + 1))
                        break
                valid_cost = curr_valid_cost
        # costs come back from the pinned per-submit ring in chunks (one host synchronisation per chunk instead of the
        # reference's one per step, vae_assoc.py:383-386); the ring holds 4096 submits
        chunk_first = submits
        def drain(upto):
            nonlocal chunk_first, avg_cost
            if upto > chunk_first:
                for cost in vae_assoc.submit_costs(chunk_first, upto - chunk_first):
                    avg_cost += cost / n_samples * batch_size
                    avg_cost_hist.append(avg_cost)
                chunk_first = upto
        for i in range(total_batch):
            if train_dev is not None:
                vae_assoc.partial_fit_indexed(train_dev, data_sets.train.next_indices(batch_size))
            else:
                batch_xs, _ = data_sets.train.next_batch(batch_size)
                slot = submits % n_slots
                slot_submit[slot] = vae_assoc.partial_fit_async(segment(batch_xs, slot))
            submits += 1
            if submits - chunk_first >= 2048:
                drain(submits)
        drain(submits)
        if epoch % display_step == 0:
            print("Epoch:", '%04d' % (epoch + 1), "cost=", "{:.9f}".format(avg_cost))
    return vae_assoc, avg_cost_hist
