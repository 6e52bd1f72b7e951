// das_tile.cu -- register-tiled full-grid power map (placeholder until the tiled kernel lands).
#include "bflk_internal.h"
namespace bflk {
int das_tile_max_span() { return -1; }
cudaError_t launch_das_tile(const TileArgs &, int, cudaStream_t, int *) { return cudaErrorNotSupported; }
}  // namespace bflk
