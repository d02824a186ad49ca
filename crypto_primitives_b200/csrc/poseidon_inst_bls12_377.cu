// Explicit instantiation of the Poseidon kernels for one field (all state widths t = 2..9).
#include "poseidon_kernels.cuh"
namespace cpb {
CPB_POS_WIDTHS(CPB_POS_INSTANTIATE, Bls12_377_Fr)
CPB_POS_INSTANTIATE_TEAM(Bls12_377_Fr)
}
