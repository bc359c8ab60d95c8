// The reference's example/thermal/main.cpp -- MainChunk::setup, MainInterface, main(), all UNCHANGED
// and read in place from the reference tree -- with the chunk base class swapped for CudaPicChunk.
//
// The four headers main.cpp includes are included first (their include guards then make main.cpp's own
// #include lines no-ops), so that the macro below renames nothing but the three uses of `PicChunk` in
// main.cpp itself: the base class of MainChunk, the inherited constructors and the call of the base
// class' setup().  Built by host/ref_binding/Makefile where the reference tree exists; nothing of the
// reference is copied into this repository.
#include "cuda_pic_chunk.hpp"

#include "nix/random.hpp"
#include "pic_application.hpp"
#include "pic_chunk.hpp"
#include "pic_diag.hpp"

#define PicChunk CudaPicChunk
#include "example/thermal/main.cpp"
