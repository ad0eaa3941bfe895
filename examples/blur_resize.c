/*
  examples/blur_resize.c -- libmagickb200 from plain C (C99): the blur + Lanczos 2x pipeline of BASELINE.json on a
  host buffer, the way a MagickCore-side caller would drive the C-ABI with a device-resident intermediate
  (include/magick_b200.h: mb200_malloc_host / mb200_malloc / mb200_upload / *_dev operators / mb200_download).

      gcc -std=c99 -Iinclude examples/blur_resize.c -Limagemagick_b200/lib -lmagickb200 \
          -Wl,-rpath,$PWD/imagemagick_b200/lib -o blur_resize && ./blur_resize 2048 2048

  Without an sm_100 GPU every operator fails with MB200_ENODEVICE (there is no CPU fallback); the program then
  reports the library version and exits 0, which is what the CPU test-suite checks.
*/
#include <stdio.h>
#include <stdlib.h>
#include "magick_b200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != MB200_OK) { \
  fprintf(stderr, "%s -> %d: %s\n", #call, rc_, mb200_last_error()); goto done; } } while (0)

int main(int argc, char **argv)
{
  const size_t w = argc > 1 ? (size_t) atol(argv[1]) : 1024, h = argc > 2 ? (size_t) atol(argv[2]) : 1024;
  const size_t in_bytes = w * h * 4 * sizeof(float), out_bytes = (w / 2) * (h / 2) * 4 * sizeof(float);
  void *h_in = NULL, *h_out = NULL, *d_in = NULL, *d_blur = NULL, *d_out = NULL;
  size_t i;
  int status = 1;

  printf("%s, %d usable device(s)\n", mb200_version(), mb200_device_count());
  if (mb200_device_count() <= 0) return 0;                 /* nothing to run on: not an error for this example */
  CHECK(mb200_set_device(0));
  CHECK(mb200_malloc_host(&h_in, in_bytes));
  CHECK(mb200_malloc_host(&h_out, out_bytes));
  CHECK(mb200_malloc(&d_in, in_bytes));
  CHECK(mb200_malloc(&d_blur, in_bytes));
  CHECK(mb200_malloc(&d_out, out_bytes));
  for (i = 0; i < w * h * 4; i++) ((float *) h_in)[i] = (float) ((i * 2654435761u) % 65536u);   /* RGBA Quantum */
  CHECK(mb200_upload(d_in, h_in, in_bytes, NULL));
  CHECK(mb200_blur_image_dev((const float *) d_in, (float *) d_blur, w, h, 4, 0.0, 4.0, NULL));       /* BlurImage(0,4) */
  CHECK(mb200_resize_image_dev((const float *) d_blur, w, h, 4, (float *) d_out, w / 2, h / 2,
                               MB200_LanczosFilter, NULL));                                           /* ResizeImage 50% */
  CHECK(mb200_download(h_out, d_out, out_bytes, NULL));
  CHECK(mb200_synchronize(NULL));
  printf("out[0..3] = %g %g %g %g (%llu kernel launches)\n", ((float *) h_out)[0], ((float *) h_out)[1],
         ((float *) h_out)[2], ((float *) h_out)[3], mb200_launch_count());
  status = 0;
done:
  if (d_out) mb200_free(d_out);
  if (d_blur) mb200_free(d_blur);
  if (d_in) mb200_free(d_in);
  if (h_out) mb200_free_host(h_out);
  if (h_in) mb200_free_host(h_in);
  return status;
}
