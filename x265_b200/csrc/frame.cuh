// x265_b200/csrc/frame.cuh -- frame-level fused kernels (lookahead + CTU analysis); see DESIGN.md.
#pragma once
#include "common.cuh"
