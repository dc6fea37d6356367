// oracle/refdev/prelude.h -- host stand-ins for the __constant__ symbols and texture references that
// Algorithm/morph.cu:24-30,592, upsample.cu:7 and render.cu:9-11 declare at file scope.  Included after the reference's
// own headers (util/dmath.h, util/linalg.h, stencils.h, Pyramid.h) and before the extracted device code.
// TEST INFRASTRUCTURE ONLY.
#pragma once
static KernParameters c_params;                       // morph.cu:592
static rod::Matrix<fmat5, 5, 5> c_tps_data;           // morph.cu:24
static rod::Matrix<imat3, 5, 5> c_improvmask;         // morph.cu:25
static imat3 c_improvmask_offset;                     // morph.cu:26
static rod::Matrix<imat5, 5, 5> c_iomask;             // morph.cu:27
static RefTex<float> tex_img0, tex_img1;              // morph.cu:29
static RefTex<float2> tex_f0, tex_f1, tex_v;          // morph.cu:30, upsample.cu:7
static RefTex<float2> tex_vector, tex_qpath;          // render.cu:9
static RefTex<float4> tex_ext0, tex_ext1;             // render.cu:11
