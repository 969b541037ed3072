// Persistent megakernel (placeholder until the state-machine kernel lands): launches the tile
// kernel with the wide traversal.
#pragma once
#include "pt_kernels.cuh"
#include "pt_wide.cuh"

namespace pt {

struct MegaState { unsigned int next_tile; unsigned int pad[3]; };

inline int launch_mega(const Scene& sc, const RenderJob& job, MegaState*, int, cudaStream_t stream)
{
    const int tiles = ((job.w + TILE_W - 1) / TILE_W) * ((job.h + TILE_H - 1) / TILE_H);
    render_tiles_kernel<WideTrav, false><<<tiles, TILE_THREADS, 0, stream>>>(sc, job, nullptr);
    return 1;
}

} // namespace pt
