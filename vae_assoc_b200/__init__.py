"""vae_assoc_b200 -- B200-native associated-VAE train step behind the reference's model surface.

    from vae_assoc_b200 import vae_assoc            # AssocVariationalAutoEncoder, train  (reference: vae_assoc.py)
    from vae_assoc_b200 import dataset              # DataSet, construct_datasets          (reference: dataset.py)
    from vae_assoc_b200 import tf_shim as tf        # name tokens the unchanged callers touch

The arithmetic lives in libvaeassoc.so (include/vaeassoc.h, csrc/*.cu); build it with
`python -m vae_assoc_b200.build`.
"""
__all__ = ["vae_assoc", "dataset", "tf_shim", "checkpoint", "build"]
