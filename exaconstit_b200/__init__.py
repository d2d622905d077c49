"""exaconstit_b200 -- B200-native (sm_100a) hot path of ExaConstit: fused crystal-plasticity
material update, matrix-free PA/EA operator apply, diagonal and residual, behind a C ABI
(include/exab200.h).  See DESIGN.md."""
__version__ = "0.1.0"
