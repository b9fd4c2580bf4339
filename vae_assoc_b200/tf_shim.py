"""The TensorFlow *names* the reference's unchanged callers touch around the model (tokens only, no arithmetic):
tf.nn.relu / tf.nn.softplus as activation selectors (vae_assoc.py:26,502; baxter_vae_assoc_writer.py:100),
tf.set_random_seed (vae_assoc_ujichar_img_jnt.py:19), tf.reset_default_graph (:94; vae_assoc_model_viewer.py:162)
and tf.all_variables (baxter_vae_assoc_writer.py:599).  `import vae_assoc_b200.tf_shim as tf` keeps those call
sites textually unchanged."""
import weakref

from . import vae_assoc as _va

_seed = [0]
_models = []          # weak references: registering a model must not keep its device memory alive


class nn(object):
    relu = staticmethod(_va.relu)
    softplus = staticmethod(_va.softplus)
    sigmoid = staticmethod(lambda x: (_ for _ in ()).throw(RuntimeError("selector token")))


def set_random_seed(seed):
    _seed[0] = int(seed)


def get_random_seed():
    return _seed[0]


def _live():
    _models[:] = [r for r in _models if r() is not None and getattr(r(), "_h", None)]
    return [r() for r in _models]


def reset_default_graph():
    """The reference releases the previous graph's resources here; close the registered models."""
    for m in _live():
        m.close()
    del _models[:]


def register(model):
    """Called by AssocVariationalAutoEncoder.__init__ (the reference's constructor adds its variables to the default graph)."""
    _models.append(weakref.ref(model))
    return model


def all_variables():
    """Names of every variable TF would list: parameters, their two Adam slots, and the two beta powers
    (86 for the reference's two dense modalities, cf. baxter_vae_assoc_writer.py:599)."""
    out = []
    for m in _live():
        names = m.variable_names()
        out += names + [n + "/Adam" for n in names] + [n + "/Adam_1" for n in names] + ["beta1_power", "beta2_power"]
    return out
