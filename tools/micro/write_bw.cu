// microbenchmark: pure streaming-write bandwidth (what a results-only kernel can reach) vs copy bandwidth
#include <cstdio>
#include <cuda_runtime.h>
__global__ void write_kernel(double2 *p, size_t n, double v)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(p + i, make_double2(v, v));
}
__global__ void copy_kernel(double2 *d, const double2 *s, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __ldcs(s + i);
}
// rows of 240,000 B written in 1 KB pieces by different CTAs, like the mesh kernel's epilogue (64 rows x 1 KB per CTA step)
__global__ void tile_write_kernel(double2 *p, size_t rows, size_t row_d2, double v)
{
	const size_t n_chunks = row_d2 / 64, tiles = rows / 64;
	for (size_t item = blockIdx.x; item < tiles * n_chunks; item += gridDim.x) {
		const size_t chunk = item / tiles, tile = item % tiles;
		for (int r = threadIdx.x / 64; r < 64; r += blockDim.x / 64) __stcs(p + (tile * 64 + r) * row_d2 + chunk * 64 + (threadIdx.x % 64), make_double2(v, v));
	}
}
int main()
{
	const size_t bytes = (size_t)48 << 30, n = bytes / 16;
	double2 *a, *b;
	cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float ms;
	for (int rep = 0; rep < 2; rep++) {
		cudaEventRecord(e0); write_kernel<<<148 * 16, 256>>>(a, n, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1); printf("write   %.0f GB/s\n", bytes / ms / 1e6);
		cudaEventRecord(e0); cudaMemsetAsync(a, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1); printf("memset  %.0f GB/s\n", bytes / ms / 1e6);
		cudaEventRecord(e0); copy_kernel<<<148 * 16, 256>>>(b, a, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1); printf("copy    %.0f GB/s (read+write)\n", 2.0 * bytes / ms / 1e6);
		const size_t row_d2 = 15000, rows = (bytes / 16 / row_d2) / 64 * 64;
		cudaEventRecord(e0); tile_write_kernel<<<148 * 2, 256>>>(a, rows, row_d2 / 64 * 64, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1); printf("tiled   %.0f GB/s (64 rows x 1 KB per CTA step, 296 CTAs)\n", rows * (row_d2 / 64 * 64) * 16.0 / ms / 1e6);
		cudaEventRecord(e0); tile_write_kernel<<<148 * 8, 256>>>(a, rows, row_d2 / 64 * 64, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
		cudaEventElapsedTime(&ms, e0, e1); printf("tiled   %.0f GB/s (1184 CTAs)\n", rows * (row_d2 / 64 * 64) * 16.0 / ms / 1e6);
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
