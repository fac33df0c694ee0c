// explicit instantiations of the tile kernels (tile.cuh) for dim 1, interp order 2
#define PHB_TILE_INSTANTIATE
#include "tile.cuh"
namespace phb
{
template int run_tile<1, 2, true>(phb_ctx*, TileMode, int, const PushParams<1>&, const DepositParams<1>&,
                                  const TileRecords&, const KeySpace<1>&, TileParams<1>&);
template int run_tile<1, 2, false>(phb_ctx*, TileMode, int, const PushParams<1>&, const DepositParams<1>&,
                                   const TileRecords&, const KeySpace<1>&, TileParams<1>&);
} // namespace phb
