/* TEST INFRASTRUCTURE. Forced-included to build the oracle for BASELINE.json configs[4]
 * ("motion-blur-heavy frame range at 4x config.hh spp"): shipped TESTING resolution and bounce
 * count (config.hh:14-18) with SAMPLES_PER_PIXEL 1024 -> 128 subframes (scene.cc:648-650). */
#include "config.hh"
#undef SAMPLES_PER_PIXEL
#define SAMPLES_PER_PIXEL 1024
