// oracle/refdev/stubs/pyramid.h -- the reference includes "pyramid.h" (stencils.cpp:6, upsample.cu:5) but the file is
// called Pyramid.h (Windows file systems are case-insensitive): forward to the reference's own header.
#pragma once
#include "Pyramid.h"
