// pharmaco_inst.cu -- the instantiations of pharmaco_kernel<N>, N = number of compartments (2 + peripheral + transit), in a
// translation unit of their own so that they compile next to the other units of libbcm3b200.so.
#include <cuda_runtime.h>

#include "pharmaco_kernel.cuh"

namespace bcm3b200 {

int launch_pharmaco(int N, dim3 grid, int block, cudaStream_t stream, const PhArgs& a)
{
	switch (N) {
	case 2: pharmaco_kernel<2><<<grid, block, 0, stream>>>(a); break;
	case 3: pharmaco_kernel<3><<<grid, block, 0, stream>>>(a); break;
	case 4: pharmaco_kernel<4><<<grid, block, 0, stream>>>(a); break;
	case 5: pharmaco_kernel<5><<<grid, block, 0, stream>>>(a); break;
	case 6: pharmaco_kernel<6><<<grid, block, 0, stream>>>(a); break;
	case 7: pharmaco_kernel<7><<<grid, block, 0, stream>>>(a); break;
	case 8: pharmaco_kernel<8><<<grid, block, 0, stream>>>(a); break;
	default: return (int)cudaErrorInvalidValue;
	}
	return (int)cudaGetLastError();
}

} // namespace bcm3b200
