// Force-included (-include) into the reference's traj2diffusion.cpp once the bodies of image() (:275-340) and slope()
// (:342-355) have been cut out and its MSD accumulation loop (:530-541) replaced by ONE line calling
// acfb200_patch_msd at build time.  See traj2diffusion_acfb200.cpp.
#pragma once
void image(double **crd, int nframes, double L);
double slope(double *msd, int n);
void acfb200_patch_msd(double **crd, int nframes, int stride, int max_s, double *msd, int *msd_counter);
