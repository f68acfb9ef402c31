// one (amplitude type, threads per CTA) instantiation of the TMA tile-pass kernel; see tile_launch.hpp
#define CB200_T float
#define CB200_NT 128
#include "tile_launch_inst.cuh"
