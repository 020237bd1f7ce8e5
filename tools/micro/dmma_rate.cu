// microbenchmark: FP64 throughput of DFMA vs mma.sync.m8n8k4.f64 (DMMA) on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double *out, int iters)
{
	double a[16], x = threadIdx.x * 1e-3, y = 1.0000001;
	for (int i = 0; i < 16; i++) a[i] = i;
	for (int it = 0; it < iters; it++)
#pragma unroll
		for (int i = 0; i < 16; i++) a[i] = fma(a[i], y, x);
	double s = 0;
	for (int i = 0; i < 16; i++) s += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_kernel(double *out, int iters)
{
	double c[8][2], a = threadIdx.x * 1e-3, b = 1.0000001;
	for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = i;
	for (int it = 0; it < iters; it++)
#pragma unroll
		for (int i = 0; i < 8; i++)
			asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
	double s = 0;
	for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
	double *out;
	cudaMalloc(&out, 148 * 8 * 256 * 8);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int blocks_per_sm : {1, 2, 4}) {
		const int grid = 148 * blocks_per_sm, iters = 20000;
		float ms;
		dfma_kernel<<<grid, 256>>>(out, 100);
		cudaEventRecord(e0); dfma_kernel<<<grid, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1);
		printf("DFMA  %d x256 thr/SM: %.2f TFLOP/s\n", blocks_per_sm, 2.0 * grid * 256 * 16.0 * iters / ms / 1e9);
		dmma_kernel<<<grid, 256>>>(out, 100);
		cudaEventRecord(e0); dmma_kernel<<<grid, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1);
		printf("DMMA  %d x256 thr/SM: %.2f TFLOP/s\n", blocks_per_sm, 2.0 * grid * 8 * 8.0 * 256 * iters / ms / 1e9);  // 8 warps x 8 mma x 256 FMA
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
