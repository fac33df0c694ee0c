// explicit instantiations of the tile kernels (tile.cuh) for dim 1, interp order 1
#define PHB_TILE_INSTANTIATE
#include "tile.cuh"
namespace phb
{
template int run_tile<1, 1, true>(phb_ctx*, TileMode, int, const PushParams<1>&, const DepositParams<1>&,
                                  const TileRecords&, const KeySpace<1>&, TileParams<1>&);
template int run_tile<1, 1, false>(phb_ctx*, TileMode, int, const PushParams<1>&, const DepositParams<1>&,
                                   const TileRecords&, const KeySpace<1>&, TileParams<1>&);
} // namespace phb
