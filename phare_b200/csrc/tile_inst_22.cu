// explicit instantiations of the tile kernels (tile.cuh) for dim 2, interp order 2
#define PHB_TILE_INSTANTIATE
#include "tile.cuh"
namespace phb
{
template int run_tile<2, 2, true>(phb_ctx*, TileMode, int, const PushParams<2>&, const DepositParams<2>&,
                                  const TileRecords&, const KeySpace<2>&, TileParams<2>&);
template int run_tile<2, 2, false>(phb_ctx*, TileMode, int, const PushParams<2>&, const DepositParams<2>&,
                                   const TileRecords&, const KeySpace<2>&, TileParams<2>&);
} // namespace phb
