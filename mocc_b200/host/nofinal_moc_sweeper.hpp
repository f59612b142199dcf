// nofinal_moc_sweeper.hpp -- forced in front of EVERY translation unit of the plugin build of MOCC
// (mocc_b200/host/Makefile, reference sources included).
//
// moc::MoCSweeper declares create_source / initialize / set_pin_flux_1g `override final`
// (src/sweepers/moc/moc_sweeper.hpp:46-95), so a subclass cannot hand MOCC's FixedSourceSolver a Source that leaves the
// fission and in-scatter source to the device, nor notice that the host has rewritten the flux
// (CMFD prolongation). The maintainer-side change is the word `final` on those three lines; this build
// emulates it without touching the reference: the keyword is defined away while that one header (and what it pulls
// in) is read. Every translation unit sees the same class, so no call is devirtualised against the override.
#pragma once
#define final
#include "sweepers/moc/moc_sweeper.hpp"
#undef final
