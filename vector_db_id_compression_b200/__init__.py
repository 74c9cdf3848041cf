"""vector_db_id_compression_b200 -- B200-native (sm_100a CUDA) ROC / Elias-Fano id codec behind the plugin
surface of facebookresearch/vector_db_id_compression.

    capi             thin ctypes front-end of the C ABI (include/idcodec.h, libidcodec.so)
    custom_invlists  host-side mirror of the reference's `custom_invlists` SWIG module (IVF inverted lists)
    altid            host-side mirror of the reference's `altid` SWIG module (NSG adjacency)
    sharding         list sharding across GPUs (LPT partition, scatter raw ids / gather blobs)
    workloads        synthetic workloads of BASELINE.json

There is no CPU fallback: importing `capi` objects works anywhere, creating a Context needs a CUDA device.
"""
__version__ = "0.1.0"
