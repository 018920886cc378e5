// One translation unit per edge count: nvcc -DNB_E=<E> -c dune_inst.cu -o dune_e<E>.o
#ifndef NB_E
#error "compile with -DNB_E=<edge count>"
#endif
#include "dune_launch.cuh"

namespace nb {
template int launch_dune_e<NB_E>(const DuneParams&, int, int, cudaStream_t, char*, size_t);
}
