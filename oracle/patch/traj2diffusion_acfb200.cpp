// The traj2diffusion bindings of INTEGRATION.md section 2 as a translation unit: image() (traj2diffusion.cpp:275-340),
// slope() (:342-355) and the MSD accumulation loop of main() (:530-541) answered by libacf_b200.so; options, readers,
// the divide and print loop (:543-547), D and its three output lines stay the reference's own code.
// oracle/Makefile links this with the cut reference source into oracle/_ref/traj2diffusion_patched.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "acf_b200.h"

namespace {
void die()
{
  std::cerr << acfb200_last_error() << std::endl;
  exit(2);
}
// the reference keeps crd[frame][0..2]; the library takes one array per coordinate
void gather(double **crd, int n, std::vector<double> c[3])
{
  for (int a = 0; a < 3; ++a) {
    c[a].resize(n);
    for (int i = 0; i < n; ++i) c[a][i] = crd[i][a];
  }
}
}  // namespace

void image(double **crd, int nframes, double L)
{
  if (nframes < 1) return;  // the reference's loops do not run
  std::vector<double> c[3];
  gather(crd, nframes, c);
  if (acfb200_image(c[0].data(), c[1].data(), c[2].data(), nframes, L) != ACFB200_OK) die();
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < nframes; ++i) crd[i][a] = c[a][i];
}

double slope(double *msd, int n)
{
  double s = 0.0, sum = 0.0;
  if (n < 1) {
    s = sum / n;  // what the reference's empty loop leaves (:354)
  } else if (acfb200_msd_slope(msd, n, &s, &sum) != ACFB200_OK) {
    die();
  }
  printf("#sum = %lf\n", sum);  // (:353)
  return s;
}

// Stands in for the loop `for(i=0;i<nframes;i+=stride) for(j<max_s) { msd[j]+=calcmsd(crd[i],crd[i+j]); ++msd_counter[j]; }`.
// The reference divides msd[j] by msd_counter[j] afterwards (:545); the library returns the quotient, so the counter is
// handed back as 1 where there were origins and 0 where there were none (msd 0.0: the reference's 0.0/0).
void acfb200_patch_msd(double **crd, int nframes, int stride, int max_s, double *msd, int *msd_counter)
{
  if (nframes < 1 || max_s < 1) return;  // msd and msd_counter were zeroed by the caller (:527-528)
  std::vector<double> c[3];
  gather(crd, nframes, c);
  std::vector<int32_t> cnt(max_s);
  if (acfb200_msd(c[0].data(), c[1].data(), c[2].data(), nframes, stride, max_s, msd, cnt.data()) != ACFB200_OK) die();
  for (int j = 0; j < max_s; ++j) {
    msd_counter[j] = cnt[j] > 0 ? 1 : 0;
    if (cnt[j] == 0) msd[j] = 0.0;
  }
}
