/* Proves the boundary is a C ABI: include/asuna_b200.h parses as C11 (gcc -std=c11 -Wall -Werror -pedantic), every entry
 * point it declares links against libasuna_b200.so, and the wire structs have the sizes of the reference's
 * src/shared headers.  Built and run by tests/test_host_and_abi.py; makes no CUDA call (asuna_abi_sizes only). */
#include <stdint.h>
#include <stdio.h>

#include "asuna_b200.h"

#define TAKE(f) sink((const void*)(uintptr_t)(f))
static int n_syms;
static void sink(const void* p) { n_syms += p != 0; }

int main(void) {
  uint32_t sz[6];
  /* every symbol of the header: taking its address forces the linker to resolve it */
  TAKE(asuna_abi_sizes); TAKE(asuna_create); TAKE(asuna_destroy); TAKE(asuna_last_error); TAKE(asuna_set_film);
  TAKE(asuna_add_texture); TAKE(asuna_set_envmap); TAKE(asuna_add_mesh); TAKE(asuna_add_material); TAKE(asuna_set_lights);
  TAKE(asuna_add_instance); TAKE(asuna_build_accel); TAKE(asuna_set_camera); TAKE(asuna_set_sunsky); TAKE(asuna_set_state);
  TAKE(asuna_reset_frame); TAKE(asuna_render_frames); TAKE(asuna_set_partition); TAKE(asuna_sync); TAKE(asuna_read_channel); TAKE(asuna_read_channel_async); TAKE(asuna_wait_reads);
  TAKE(asuna_export_partial); TAKE(asuna_post_process); TAKE(asuna_import_partial); TAKE(asuna_host_alloc); TAKE(asuna_host_free);
  TAKE(asuna_channel_device_ptr); TAKE(asuna_stream_handle); TAKE(asuna_set_counting); TAKE(asuna_set_profiling);
  TAKE(asuna_get_stats); TAKE(asuna_reset_stats); TAKE(asuna_trace_primary); TAKE(asuna_trace_rays); TAKE(asuna_occlusion_rays);
  TAKE(asuna_accel_stats);
  asuna_abi_sizes(sz);
  printf("%d %u %u %u %u %u %u %u %u %u %u %u %u\n", n_syms, sz[0], sz[1], sz[2], sz[3], sz[4], sz[5],
         (unsigned)sizeof(AsunaVertex), (unsigned)sizeof(AsunaMaterial), (unsigned)sizeof(AsunaLight),
         (unsigned)sizeof(AsunaCamera), (unsigned)sizeof(AsunaState), (unsigned)sizeof(AsunaSunSky));
  return !(sz[0] == sizeof(AsunaVertex) && sz[1] == sizeof(AsunaMaterial) && sz[2] == sizeof(AsunaLight) &&
           sz[3] == sizeof(AsunaCamera) && sz[4] == sizeof(AsunaState) && sz[5] == sizeof(AsunaSunSky) &&
           sz[0] == 44 && sz[1] == 132 && sz[2] == 76 && sz[3] == 224 && sz[4] == 84 && sz[5] == 96);
}
