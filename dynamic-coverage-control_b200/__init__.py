"""dcc_b200 — B200-native (sm_100a) hot path of zhaozijie2022/dynamic-coverage-control.

The directory is named `dynamic-coverage-control_b200` (not importable as such); the sibling `dcc_b200/`
shim package makes it importable as `dcc_b200`.

Product code only: hand-written CUDA behind the C ABI in include/dcc_b200.h (`libdcc_b200.so`, built by
`dcc_b200.build`), plus the thin Python host that mirrors the reference's plugin interface
(`envs.make_env.make_env(cfg)` -> vec-env with reset()/step()).  Nothing here imports `oracle/`.
"""
__version__ = "0.1.0"
