// poppk_inst_transit.cu -- the transit instances of poppk_kernel (see poppk_kernel.cuh).
#include <cuda_runtime.h>

#include "poppk_kernel.cuh"

namespace bcm3b200 {

template <class Model, bool DIAG, int STRIDE>
static int launch_one(dim3 grid, int block, size_t smem_bytes, cudaStream_t stream, const PkArgs& a)
{
	if (smem_bytes > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(poppk_kernel<Model, DIAG, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
		if (e != cudaSuccess) return (int)e;
	}
	poppk_kernel<Model, DIAG, STRIDE><<<grid, block, smem_bytes, stream>>>(a);
	return (int)cudaGetLastError();
}

template <class Model>
static int launch_model(bool diagnostics, int stride, dim3 grid, int block, size_t smem_bytes, cudaStream_t stream, const PkArgs& a)
{
	constexpr int SMALL = BCM3_POPPK_STRIDE_SMALL, BIG = BCM3_POPPK_STRIDE_BIG;
	if (diagnostics) return stride == SMALL ? launch_one<Model, true, SMALL>(grid, block, smem_bytes, stream, a) : launch_one<Model, true, BIG>(grid, block, smem_bytes, stream, a);
	return stride == SMALL ? launch_one<Model, false, SMALL>(grid, block, smem_bytes, stream, a) : launch_one<Model, false, BIG>(grid, block, smem_bytes, stream, a);
}

int launch_poppk_transit(bool two, bool diagnostics, int stride, dim3 grid, int block, size_t smem_bytes, cudaStream_t stream, const PkArgs& a)
{
	if (two) return launch_model<PkTwoTransitModel>(diagnostics, stride, grid, block, smem_bytes, stream, a);
	return launch_model<PkOneTransitModel>(diagnostics, stride, grid, block, smem_bytes, stream, a);
}

} // namespace bcm3b200
