// explicit instantiations of the tile kernels (tile.cuh) for dim 3, interp order 1
#define PHB_TILE_INSTANTIATE
#include "tile.cuh"
namespace phb
{
template int run_tile<3, 1, true>(phb_ctx*, TileMode, int, const PushParams<3>&, const DepositParams<3>&,
                                  const TileRecords&, const KeySpace<3>&, TileParams<3>&);
template int run_tile<3, 1, false>(phb_ctx*, TileMode, int, const PushParams<3>&, const DepositParams<3>&,
                                   const TileRecords&, const KeySpace<3>&, TileParams<3>&);
} // namespace phb
