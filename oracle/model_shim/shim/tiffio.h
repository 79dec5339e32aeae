// Hand-declared prototypes of the libtiff 4 entry points src/model/Grid.cpp:27-79 calls.
// No libtiff headers exist in this image; the symbols resolve against Pillow's bundled
// libtiff (pillow.libs/libtiff-*.so), i.e. the REAL TIFFReadRGBAImage.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
extern "C" {
typedef struct tiff TIFF;
TIFF* TIFFOpen(const char*, const char*);
void TIFFClose(TIFF*);
int TIFFGetField(TIFF*, uint32_t, ...);
int TIFFReadDirectory(TIFF*);
int TIFFSetDirectory(TIFF*, uint32_t);
int TIFFReadRGBAImage(TIFF*, uint32_t, uint32_t, uint32_t*, int = 0);
}
#define TIFFTAG_IMAGEWIDTH 256
#define TIFFTAG_IMAGELENGTH 257
