/* TEST INFRASTRUCTURE. Forced-included (-include) before any reference source to build the oracle at
 * the PRODUCTION settings of config.hh:21-25 (TESTING undefined) without editing the reference:
 * the reference header is pulled in first (its include guard then makes later includes no-ops)
 * and the five size macros are replaced. Common settings (config.hh:28-42) are left as they are. */
#include "config.hh"
#undef TESTING
#undef IMAGE_WIDTH
#undef IMAGE_HEIGHT
#undef SAMPLES_PER_PIXEL
#undef MAX_BOUNCES
#define IMAGE_WIDTH 1920
#define IMAGE_HEIGHT 1080
#define SAMPLES_PER_PIXEL 1024
#define MAX_BOUNCES 5
