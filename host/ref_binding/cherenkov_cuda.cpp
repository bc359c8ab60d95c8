// The reference's example/cherenkov/main.cpp -- MainChunk::setup, MainInterface, main(), all UNCHANGED and read in
// place from the reference tree -- with the chunk base class swapped for CudaPicChunk (see thermal_cuda.cpp).
#include "cuda_pic_chunk.hpp"

#include "nix/random.hpp"
#include "pic_application.hpp"
#include "pic_chunk.hpp"
#include "pic_diag.hpp"

#define PicChunk CudaPicChunk
#include "example/cherenkov/main.cpp"
