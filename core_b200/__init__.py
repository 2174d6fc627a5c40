"""core_b200 -- B200-native MeshAdapt marking / quality sweep (drop-in for that path of SCOREC/core).

`Part` (sweep.py) mirrors the reference interface over the C ABI of include/mag.h;
`boxmesh` generates Kuhn boxes in apf::makeMdsBox order; `fields` holds the synthetic
benchmark size fields.  CUDA only: there is no CPU fallback."""
from .sweep import (Part, MagError, MAXLENGTH, MINLENGTH, GOOD_QUALITY_3D, GOOD_QUALITY_2D,
                    SPLIT, DONT_SPLIT, COLLAPSE, DONT_COLLAPSE, CHECKED, BAD_QUALITY, OK_QUALITY, DONT_SWAP, LAYER,
                    NEED_NOT_SPLIT, NEED_NOT_COLLAPSE,
                    OP_LENGTHS, OP_MARK_SPLIT, OP_MARK_COLLAPSE, OP_QUALITIES, OP_MARK_BAD, OP_LAYER_CHECK, OP_ALL, OP_LENGTH_SUM,
                    FP_STRICT, FP_FAST, FP_FAST_LISTED)
from . import boxmesh, fields
