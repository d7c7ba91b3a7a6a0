// GPU marching cubes for extract_geometry (reference: models/renderer.py:32-40 -> PyMCubes on the CPU after a 512 MiB
// device-to-host copy).  Two passes around two prefix sums:
//   classify: per grid point the crossings on the 3 edges it owns (+x, +y, +z), per cell the triangle count of its case
//   emit    : per grid point its vertices (linear interpolation on the crossed edge, grid-index coordinates like PyMCubes),
//             per cell its triangles, every corner resolved to the shared vertex of the owning grid point
// The case table is derived on the host (factored-neus_b200/mcubes.py) and passed in.  Inside: u > isovalue.
#include "fneus_common.cuh"
#include "prof.cuh"

namespace fneus {

// corner c = (x, y, z) bits; edge e joins the corners below; its owner is the grid point at the edge's lower end
__constant__ int c_mc_edge_a[12] = {0, 2, 4, 6, 0, 1, 4, 5, 0, 1, 2, 3};
__constant__ int c_mc_edge_axis[12] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2};

__device__ __forceinline__ long long mc_pt(int x, int y, int z, int ny, int nz) { return ((long long)x * ny + y) * nz + z; }

__global__ void mc_classify_kernel(const float* __restrict__ u, int nx, int ny, int nz, float iso,
                                   const int* __restrict__ tri_count, unsigned char* __restrict__ vmask,
                                   int* __restrict__ vcount, int* __restrict__ tcount) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npts = (long long)nx * ny * nz;
  if (p >= npts) return;
  const int z = (int)(p % nz);
  const long long t = p / nz;
  const int y = (int)(t % ny), x = (int)(t / ny);
  const bool in0 = u[p] > iso;
  unsigned m = 0;
  if (x + 1 < nx && (u[mc_pt(x + 1, y, z, ny, nz)] > iso) != in0) m |= 1u;
  if (y + 1 < ny && (u[mc_pt(x, y + 1, z, ny, nz)] > iso) != in0) m |= 2u;
  if (z + 1 < nz && (u[mc_pt(x, y, z + 1, ny, nz)] > iso) != in0) m |= 4u;
  vmask[p] = (unsigned char)m;
  vcount[p] = __popc(m);
  if (x + 1 < nx && y + 1 < ny && z + 1 < nz) {
    unsigned cs = 0;
#pragma unroll
    for (int c = 0; c < 8; c++)
      if (u[mc_pt(x + (c & 1), y + ((c >> 1) & 1), z + ((c >> 2) & 1), ny, nz)] > iso) cs |= 1u << c;
    tcount[((long long)x * (ny - 1) + y) * (nz - 1) + z] = tri_count[cs];
  }
}

__global__ void mc_emit_kernel(const float* __restrict__ u, int nx, int ny, int nz, float iso,
                               const int* __restrict__ tri_count, const int* __restrict__ tri_edges, int maxt,
                               const unsigned char* __restrict__ vmask, const int* __restrict__ vcount,
                               const long long* __restrict__ voff_incl, const int* __restrict__ tcount,
                               const long long* __restrict__ toff_incl, float* __restrict__ verts,
                               long long* __restrict__ tris) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npts = (long long)nx * ny * nz;
  if (p >= npts) return;
  const int z = (int)(p % nz);
  const long long t = p / nz;
  const int y = (int)(t % ny), x = (int)(t / ny);
  // ---- vertices of the edges this grid point owns ----
  const unsigned m = vmask[p];
  if (m) {
    long long v = voff_incl[p] - vcount[p];
    const float u0 = u[p];
    const int step[3] = {1, 0, 0};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (!(m & (1u << a))) continue;
      const long long q = mc_pt(x + (a == 0), y + (a == 1), z + (a == 2), ny, nz);
      const float u1 = u[q];
      const float w = __fdiv_rn(__fsub_rn(iso, u0), __fsub_rn(u1, u0));       // crossing parameter along the edge
      verts[v * 3 + 0] = (float)x + (a == 0 ? w : 0.f);
      verts[v * 3 + 1] = (float)y + (a == 1 ? w : 0.f);
      verts[v * 3 + 2] = (float)z + (a == 2 ? w : 0.f);
      v++;
    }
    (void)step;
  }
  // ---- triangles of the cell whose lower corner is this grid point ----
  if (x + 1 < nx && y + 1 < ny && z + 1 < nz) {
    const long long cell = ((long long)x * (ny - 1) + y) * (nz - 1) + z;
    const int nt = tcount[cell];
    if (nt == 0) return;
    unsigned cs = 0;
#pragma unroll
    for (int c = 0; c < 8; c++)
      if (u[mc_pt(x + (c & 1), y + ((c >> 1) & 1), z + ((c >> 2) & 1), ny, nz)] > iso) cs |= 1u << c;
    long long o = toff_incl[cell] - nt;
    for (int i = 0; i < nt * 3; i++) {
      const int e = tri_edges[cs * 3 * maxt + i];
      const int ca = c_mc_edge_a[e], ax = c_mc_edge_axis[e];
      const long long q = mc_pt(x + (ca & 1), y + ((ca >> 1) & 1), z + ((ca >> 2) & 1), ny, nz);
      const unsigned mq = vmask[q];
      // rank of axis `ax` among the owner's crossed edges
      const long long vid = voff_incl[q] - vcount[q] + __popc(mq & ((1u << ax) - 1u));
      tris[(o + i / 3) * 3 + (i % 3)] = vid;
    }
  }
}

}  // namespace fneus

using namespace fneus;

extern "C" {

int fneus_mc_classify(const float* u, int nx, int ny, int nz, float isovalue, const int* tri_count,
                      unsigned char* vmask, int* vcount, int* tcount, void* stream) {
  if (!u || !tri_count || !vmask || !vcount || !tcount) return FNEUS_ERR_NULL;
  if (nx < 1 || ny < 1 || nz < 1) return FNEUS_ERR_BAD_SHAPE;
  const long long npts = (long long)nx * ny * nz;
  prof_begin(PC_ELEMENTWISE, 0.0, (double)npts * 4.0 * 3.0, (cudaStream_t)stream);
  mc_classify_kernel<<<cdiv(npts, 256), 256, 0, (cudaStream_t)stream>>>(u, nx, ny, nz, isovalue, tri_count, vmask, vcount,
                                                                        tcount);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_mc_emit(const float* u, int nx, int ny, int nz, float isovalue, const int* tri_count, const int* tri_edges,
                  int max_tris, const unsigned char* vmask, const int* vcount, const long long* voff_incl,
                  const int* tcount, const long long* toff_incl, float* verts, long long* tris, void* stream) {
  if (!u || !tri_count || !tri_edges || !vmask || !vcount || !voff_incl || !tcount || !toff_incl || !verts || !tris)
    return FNEUS_ERR_NULL;
  if (nx < 2 || ny < 2 || nz < 2 || max_tris < 1 || max_tris > 16) return FNEUS_ERR_BAD_SHAPE;
  const long long npts = (long long)nx * ny * nz;
  prof_begin(PC_ELEMENTWISE, 0.0, (double)npts * 4.0 * 3.0, (cudaStream_t)stream);
  mc_emit_kernel<<<cdiv(npts, 256), 256, 0, (cudaStream_t)stream>>>(u, nx, ny, nz, isovalue, tri_count, tri_edges, max_tris,
                                                                    vmask, vcount, voff_incl, tcount, toff_incl, verts, tris);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"
