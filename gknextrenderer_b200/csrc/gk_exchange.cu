// gk_exchange.cu — frame-end exchange helpers of the multi-GPU compositor.
//
// The reference is single-GPU; this is the one step the tile-partitioned frame adds.  Rank r owns
// the row tiles with (row / tileRows) % tileCount == r.  At frame end
//   pack   : the owned rows of the six integrator planes the filters read (diffuse, specular,
//            albedo, normal, object id, motion = 44 B/pixel) are gathered into one contiguous
//            staging buffer                                   (this file, one launch)
//   gather : ncclAllGather over NVLink on the staging buffers (host side: torch.distributed)
//   unpack : every rank scatters all ranks' rows back into its full planes (one launch)
// after which each rank filters the whole image locally (halos need no further traffic).
// Both kernels are pure streaming copies in 4-byte words, coalesced along rows.
#include "gk_context.h"

namespace gk {

namespace {

constexpr int kExchangePlanes = 6;
const GkPlane kPlaneIds[kExchangePlanes] = {GK_PLANE_OUTPUT_DIFFUSE, GK_PLANE_OUTPUT_SPECULAR, GK_PLANE_ALBEDO, GK_PLANE_NORMAL, GK_PLANE_OBJECT_ID0, GK_PLANE_MOTION};
const uint32_t kPlaneWords[kExchangePlanes] = {2, 2, 2, 2, 1, 2}; // 4-byte words per pixel

struct ExchangeArgs {
    uint32_t* plane[kExchangePlanes];
    uint32_t words[kExchangePlanes];    // words per pixel
    uint64_t offset[kExchangePlanes];   // word offset of the plane inside one rank's staging block
    uint32_t width, height, tileRows, tileCount, blocksPerRank;
    uint64_t rankWords; // words per rank block
};

// One thread per word of the staging buffer of `ranks` ranks (pack: ranks == 1 and rank0 = own rank).
template <bool kPack>
__global__ void __launch_bounds__(256) k_exchange(ExchangeArgs A, uint32_t* __restrict__ staging, uint32_t rank0, uint32_t ranks)
{
    const uint64_t total = A.rankWords * ranks;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t rk = (uint32_t)(i / A.rankWords) + rank0;
        uint64_t w = i % A.rankWords;
        int p = kExchangePlanes - 1;
#pragma unroll
        for (int q = kExchangePlanes - 1; q > 0; --q)
            if (w < A.offset[q]) p = q - 1;
        w -= A.offset[p];
        const uint32_t rowWords = A.width * A.words[p];
        const uint32_t lrow = (uint32_t)(w / rowWords), col = (uint32_t)(w % rowWords);
        const uint32_t blk = lrow / A.tileRows, within = lrow % A.tileRows;
        const uint32_t row = (blk * A.tileCount + rk) * A.tileRows + within;
        if (row >= A.height) continue; // padding rows of the last block
        uint32_t* px = A.plane[p] + (uint64_t)row * rowWords + col;
        if (kPack) staging[i] = *px;
        else *px = staging[i];
    }
}

ExchangeArgs makeArgs(const Context& c)
{
    ExchangeArgs A;
    A.width = c.width, A.height = c.height, A.tileRows = c.tileRows, A.tileCount = c.tileCount;
    A.blocksPerRank = (c.height + c.tileRows * c.tileCount - 1) / (c.tileRows * c.tileCount);
    uint64_t off = 0;
    for (int p = 0; p < kExchangePlanes; ++p) {
        A.plane[p] = (uint32_t*)c.planes.p[kPlaneIds[p]];
        A.words[p] = kPlaneWords[p];
        A.offset[p] = off;
        off += (uint64_t)A.blocksPerRank * c.tileRows * c.width * kPlaneWords[p];
    }
    A.rankWords = off;
    return A;
}

} // namespace

size_t exchangeBytesPerRank(const Context& c) { return (size_t)makeArgs(c).rankWords * 4; }

GkStatus exchangePack(Context& c, void* dStaging)
{
    const ExchangeArgs A = makeArgs(c);
    const unsigned grid = (unsigned)std::min<uint64_t>((A.rankWords + 255) / 256, 148u * 16u);
    k_exchange<true><<<grid, 256, 0, c.stream>>>(A, (uint32_t*)dStaging, c.tileIndex, 1);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return GK_OK;
}

GkStatus exchangeUnpack(Context& c, const void* dAll)
{
    const ExchangeArgs A = makeArgs(c);
    const unsigned grid = (unsigned)std::min<uint64_t>((A.rankWords * c.tileCount + 255) / 256, 148u * 16u);
    k_exchange<false><<<grid, 256, 0, c.stream>>>(A, (uint32_t*)dAll, 0, c.tileCount);
    GK_CUDA(cudaGetLastError());
    c.stats.launches += 1;
    return GK_OK;
}

} // namespace gk
