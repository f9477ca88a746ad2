/* The C ABI of libzkb.so driven from plain C through dlopen -- what a zkb-sys crate's build would link against.
 * Test infrastructure (built and run by tests/test_bindings.py).
 *
 *   abi_dlopen <libzkb.so> <symbols.txt> <fixture.bin>
 *
 * 1. every symbol named in symbols.txt (one per line, extracted from include/zkb.h by the test) resolves;
 * 2. the host-only helpers work without a GPU (Keccak-f[1600] of the zero state, first lane of the published vector);
 * 3. zkb_init: without a device it must fail with ZKB_E_NO_DEVICE (no CPU fallback) -> prints "no-device" and exits 0;
 * 4. with a device: the byte-layout fixture -- ark-serialize compressed points (fixture.bin, written by the test from the
 *    oracle) -> zkb_points_decompress -> compare with the expected x||y Montgomery limbs -> zkb_srs_upload -> zkb_msm with
 *    the fixture's canonical scalars -> compare with the expected affine sum.
 *
 * fixture.bin: u32 curve, u32 group, u32 n, u32 words (u64 limbs per affine point), then n * words*4 compressed bytes,
 * n * words u64 expected points, n bytes expected infinity flags, n * 4 u64 canonical scalars, words u64 expected MSM
 * result, 1 byte expected identity flag. */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/zkb.h"

#define FAIL(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); return 1; } while (0)

typedef int (*init_fn)(int, zkb_ctx**);
typedef void (*destroy_fn)(zkb_ctx*);
typedef const char* (*err_fn)(zkb_ctx*);
typedef int (*decompress_fn)(zkb_ctx*, int, int, const uint8_t*, size_t, unsigned, uint64_t*, uint8_t*, uint8_t*);
typedef int (*upload_fn)(zkb_ctx*, int, int, const uint64_t*, const uint8_t*, size_t, unsigned, zkb_srs**);
typedef void (*srs_free_fn)(zkb_srs*);
typedef int (*msm_fn)(zkb_ctx*, const zkb_srs*, size_t, const uint64_t*, size_t, uint64_t*, uint8_t*);
typedef void (*keccak_fn)(uint64_t*);
typedef uint64_t (*count_fn)(zkb_ctx*);

int main(int argc, char** argv) {
  if (argc < 4) FAIL("usage: abi_dlopen <lib> <symbols> <fixture>");
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) FAIL("dlopen: %s", dlerror());

  FILE* sf = fopen(argv[2], "r");
  if (!sf) FAIL("cannot open %s", argv[2]);
  char name[256];
  int n_syms = 0;
  while (fgets(name, sizeof name, sf)) {
    name[strcspn(name, "\r\n")] = 0;
    if (!name[0]) continue;
    if (!dlsym(h, name)) FAIL("missing symbol %s", name);
    n_syms++;
  }
  fclose(sf);
  printf("symbols %d\n", n_syms);

  uint64_t st[25];
  memset(st, 0, sizeof st);
  ((keccak_fn)dlsym(h, "zkb_host_keccak_f1600"))(st);
  if (st[0] != 0xF1258F7940E1DDE7ull) FAIL("keccak-f[1600] of the zero state: lane 0 = %016llx", (unsigned long long)st[0]);
  printf("keccak ok\n");

  zkb_ctx* ctx = NULL;
  int rc = ((init_fn)dlsym(h, "zkb_init"))(0, &ctx);
  if (rc == ZKB_E_NO_DEVICE) {
    if (ctx) FAIL("zkb_init failed but returned a context");
    printf("no-device\n");
    return 0;
  }
  if (rc != ZKB_OK) FAIL("zkb_init rc=%d", rc);
  err_fn last_error = (err_fn)dlsym(h, "zkb_last_error");

  FILE* f = fopen(argv[3], "rb");
  if (!f) FAIL("cannot open %s", argv[3]);
  uint32_t hdr[4];
  if (fread(hdr, 4, 4, f) != 4) FAIL("short fixture");
  int curve = (int)hdr[0], group = (int)hdr[1];
  size_t n = hdr[2], words = hdr[3];
  size_t cbytes = words * 4;
  uint8_t* comp = malloc(n * cbytes);
  uint64_t* want_xy = malloc(n * words * 8);
  uint8_t* want_inf = malloc(n);
  uint64_t* scalars = malloc(n * 32);
  uint64_t* want_sum = malloc(words * 8);
  uint8_t want_sum_inf;
  if (fread(comp, cbytes, n, f) != n || fread(want_xy, words * 8, n, f) != n || fread(want_inf, 1, n, f) != n ||
      fread(scalars, 32, n, f) != n || fread(want_sum, 8, words, f) != words || fread(&want_sum_inf, 1, 1, f) != 1)
    FAIL("short fixture");
  fclose(f);

  uint64_t* xy = calloc(n * words, 8);
  uint8_t* inf = calloc(n, 1);
  uint8_t* status = calloc(n, 1);
  rc = ((decompress_fn)dlsym(h, "zkb_points_decompress"))(ctx, curve, group, comp, n, ZKB_DECOMPRESS_CHECK_SUBGROUP, xy, inf, status);
  if (rc != ZKB_OK) FAIL("zkb_points_decompress rc=%d: %s", rc, last_error(ctx));
  for (size_t i = 0; i < n; i++)
    if (status[i] != 0) FAIL("point %zu rejected with status %d", i, status[i]);
  if (memcmp(inf, want_inf, n)) FAIL("infinity flags differ");
  if (memcmp(xy, want_xy, n * words * 8)) FAIL("decompressed limbs differ from the expected ark-ff Montgomery layout");
  printf("decompress ok (%zu points, curve %d group %d)\n", n, curve, group);

  zkb_srs* srs = NULL;
  rc = ((upload_fn)dlsym(h, "zkb_srs_upload"))(ctx, curve, group, xy, inf, n, ZKB_SRS_PRECOMPUTE, &srs);
  if (rc != ZKB_OK) FAIL("zkb_srs_upload rc=%d: %s", rc, last_error(ctx));
  uint64_t* sum = calloc(words, 8);
  uint8_t sum_inf = 9;
  rc = ((msm_fn)dlsym(h, "zkb_msm"))(ctx, srs, 0, scalars, n, sum, &sum_inf);
  if (rc != ZKB_OK) FAIL("zkb_msm rc=%d: %s", rc, last_error(ctx));
  if (sum_inf != want_sum_inf || (!sum_inf && memcmp(sum, want_sum, words * 8))) FAIL("MSM result differs");
  if (((count_fn)dlsym(h, "zkb_launch_count"))(ctx) == 0) FAIL("no kernel was launched");
  printf("msm ok\n");

  /* error behaviour: a null SRS is an argument error, not a crash */
  rc = ((msm_fn)dlsym(h, "zkb_msm"))(ctx, NULL, 0, scalars, n, sum, &sum_inf);
  if (rc != ZKB_E_INVALID) FAIL("zkb_msm(NULL srs) rc=%d, want ZKB_E_INVALID", rc);

  ((srs_free_fn)dlsym(h, "zkb_srs_free"))(srs);
  ((destroy_fn)dlsym(h, "zkb_destroy"))(ctx);
  printf("gpu ok\n");
  return 0;
}
