"""rd_b200 — B200-native hot path of representation-disentanglement (IPMI 2021).

The package is deliberately small: `csrc/` (sm_100a CUDA kernels behind a C ABI, see include/rd_b200.h),
`lib.py` (ctypes binding that fails loudly without the built library), `ops.py` (autograd glue),
`model.py` (the reference's nn.Module classes: same names, ctor arguments and state_dict keys),
`trainer.py` (the loop body of the reference main_missing.py), `ddp.py`, `data.py`, `config.py`.
"""
__version__ = "0.1.0"
