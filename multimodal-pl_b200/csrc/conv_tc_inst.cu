// One (MODE, EPI) family of tcgen05 convolution kernels per object file: compiled several times with
// -DMMPL_TC_MODE=<0..5> -DMMPL_TC_EPI=<0..2> (Makefile), so the families build in parallel.
#include "conv_tc_impl.cuh"

namespace mmpl {
template int dispatch_tc<MMPL_TC_MODE, MMPL_TC_EPI>(const TcProblem&, cudaStream_t);
}  // namespace mmpl
