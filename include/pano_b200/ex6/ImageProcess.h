// pano_b200/ex6/ImageProcess.h -- drop-in for the reference's SECOND copy of the class, src/ex6/ImageProcess.h:75-147
// (fixed left-to-right chain, min(w,h) pyramid depth, Deriche pyramid blur, 5/6 : 1/6 luminance mix, result saved as
// <dir>result.bmp).  Same constructor ImageProcess(std::string dir, int n).  The reference seeds RANSAC with time(0)
// (src/ex6/ImageProcess.cpp:403); define PANO_B200_RANSAC_SEED before including this header for a reproducible run.
#ifndef PANO_B200_EX6_IMAGEPROCESS_H
#define PANO_B200_EX6_IMAGEPROCESS_H
#define PANO_B200_IMAGEPROCESS_EX6 1
#include "../ImageProcess.h"
#endif
