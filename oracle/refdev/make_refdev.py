#!/usr/bin/env python
"""oracle/refdev/make_refdev.py -- builds oracle/_ref/libref_devfn.so: the REFERENCE's own CUDA device code for the hot
path (Algorithm/morph.cu, upsample.cu, render.cu) and its stencils.cpp, compiled for the HOST behind the SIMT emulator of
simt.h, so tests can check the oracle's restatement against reference code (tests/test_oracle_refdev.py).

TEST INFRASTRUCTURE ONLY.  Nothing of the reference is copied into the repository: the device functions are cut out of
the reference files where they lie (by the anchors below, not by line number) into a temporary translation unit that is
deleted after compilation; only the shared library lands in oracle/_ref/ (git-ignored, travels to the GPU box).
The reference's headers (util/dmath.h, util/linalg.h, stencils.h, Pyramid.h, parameters.h) are included in place;
stubs/ only holds stand-ins for things that do not exist on this machine (OpenCV-C++: the handful of cv::Mat operations
Morph::cpu_optimize_level uses; CUDA's internal device_functions.h; the case-insensitive "pyramid.h").

    python oracle/refdev/make_refdev.py [--ref /root/reference] [--keep]
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "..", "_ref")
OUT = os.path.join(OUT_DIR, "libref_devfn.so")

# (file, regex of the first line taken, regex of the first line NOT taken, regex of trailing lines to drop or None
#  [, substitutions [, prefix, suffix]])
SPANS = [
    # Morph::cpu_optimize_level                                      morph.cu:419-590.  The member function becomes a free function
    # (ONE substitution, on its signature line: the body is the reference's text; m_params is passed in) because Morph's
    # constructor and the rest of the class live in parts of morph.cu that need the GPU.  cv::Mat = stubs/opencv2/mat_stub.h.
    ("Algorithm/morph.cu", r"^void Morph::cpu_optimize_level\(", r"^__constant__ KernParameters c_params;", r"^\s*$",
     [(r"^void Morph::cpu_optimize_level\(PyramidLevel &lvl,PyramidLevel &lv0\)", "static void ref_cpu_optimize_level(Parameters &m_params, PyramidLevel &lvl, PyramidLevel &lv0)")]),
    # the host UI splat at the end of Morph::initialize_level          morph.cu:341-388.  Only the tail of that function can be
    # compiled here (its head launches kernels); the cut lines are wrapped (prefix / suffix below, not reference text) into a
    # function with the three names the body uses.  fabs / floor / ceil of a float are the float overloads, as under MSVC.
    ("Algorithm/morph.cu", r"^\s*// initialize ui data in cpu", r"^void Morph::clear_level", r"^\s*$|^\}\s*$", [],
     "namespace ref_host { using std::fabs; using std::floor; using std::ceil;\nstatic void ref_ui_splat(Parameters &m_params, PyramidLevel &lvl, PyramidLevel &lv0)\n{\n", "}\n}\n"),
    # CQuadraticPath::optimize                                         QuadraticPath.cpp:24-223: blended Jacobians, the right-hand
    # sides and the CSR matrix of the two Poisson systems.  The class is a QThread; its optimize() becomes a member of a plain
    # struct with the same member names (ONE substitution, on the signature line), whose cudaSolver (cuSPARSE / cuBLAS in the
    # reference, QuadraticPath.cpp:225-318) records the system it is handed.  sqrt of a float is the float overload, as under MSVC.
    ("Algorithm/QuadraticPath.cpp", r"^void CQuadraticPath::optimize\(\)", r"^void CQuadraticPath::cudaSolver\(", r"^\s*$",
     [(r"^void CQuadraticPath::optimize\(\)", "void RefQPath::optimize()")],
     "namespace ref_host { using std::sqrt; using cv::Vec2f;\nstruct RefQPath { std::vector<cv::Mat> &_vector, &_qpath; int times, rows, cols;\n"
     "    void cudaSolver(float *A, int *rowindex, int *columns, int N, int nz, float *B, float *X); void optimize(); };\n", "}\n"),
    # CMatchingThread::Resize + BiLinear                                MatchingThread.cpp:86-136: the spatial resample of update_result.
    # The class is a QThread: the two member definitions become members of a plain struct (one substitution each, class name only).
    ("Algorithm/MatchingThread.cpp", r"^void CMatchingThread::Resize\(", r"^void CMatchingThread::run\(\)", r"^\s*$",
     [(r"^void CMatchingThread::Resize\(Mat& src,Mat& dst\)", "void RefMatch::Resize(cv::Mat& src, cv::Mat& dst)"),
      (r"^inline T CMatchingThread::BiLinear\(", "inline T RefMatch::BiLinear(")],
     "namespace ref_host { using std::floor; using std::ceil; using cv::Vec2f;\nstruct RefMatch { void Resize(cv::Mat &src, cv::Mat &dst); template <class T> T BiLinear(cv::Mat &img, float2 p); };\n", "}\n"),
    # the prolongation of `upsample` (upsample.cu:259-285): internal_vector_to_image (pyramid.cu:627-644, its `template <class T>`
    # line re-added as prefix), rod::kernel_upsample (imgop_upsample.cu:12-31, inside namespace rod as in the reference) and
    # conv_to_block_of_arrays (upsample.cu:9-26)
    ("Algorithm/pyramid.cu", r"^__global__ void internal_vector_to_image\(rod::dimage_ptr<T> res,", r"^template <class T>\s*$", r"^\s*$", [],
     "template <class T>\n", ""),
    ("include/util/imgop_upsample.cu", r"^const int BW = 32,", r"^template <class T, int C>\s*$", r"^\s*$", [], "namespace rod {\n", "}\n"),
    ("Algorithm/upsample.cu", r"^__global__ void conv_to_block_of_arrays\(", r"^__global__ void temp_ref\(", r"^\s*$"),
    # the temporal flow composition of Pyramid::build (pyramid.cu:406-441: a block in the middle of that function, wrapped into a
    # member of a plain struct with the locals it uses as parameters) and Pyramid::BiLinear (488-523, class name substituted)
    ("Algorithm/pyramid.cu", r"^\t\t\tif\(factor_t>1\)\s*$", r"^\t\tfor\(int t=0;t<d;t\+\+\)\s*$", r"^\s*$", [],
     "namespace ref_host { using std::floor; using std::ceil;\nstruct RefPyr { template <class T> T BiLinear(cv::Mat &img, float2 p);\n"
     "    void compose(std::vector<cv::Mat> &forw0, std::vector<cv::Mat> &forw1, std::vector<cv::Mat> &back0, std::vector<cv::Mat> &back1, int d, int h, int w, int prev_d, int factor_t); };\n"
     "void RefPyr::compose(std::vector<cv::Mat> &forw0, std::vector<cv::Mat> &forw1, std::vector<cv::Mat> &back0, std::vector<cv::Mat> &back1, int d, int h, int w, int prev_d, int factor_t)\n{\n",
     "}\n}\n"),
    ("Algorithm/pyramid.cu", r"^inline T Pyramid::BiLinear\(", r"^PyramidLevel &Pyramid::append_new\(", r"^\s*$",
     [(r"^inline T Pyramid::BiLinear\(", "inline T RefPyr::BiLinear(")], "namespace ref_host { using cv::Vec2f;\ntemplate <class T>\n", "}\n"),
    # the level schedule of Pyramid::build: the block that derives the number of levels (pyramid.cu:222-234) and the three lines
    # that shrink w / h / d at the end of the level loop (463-465), stitched into a function whose loop skeleton (prefix /
    # suffix strings below) records the sizes append_new() would be called with.  log2 / sqrt / ceil of a float are the float
    # overloads, as under MSVC.
    ("Algorithm/pyramid.cu", r"^\tfloat decres_fa=\(float\)\(w\*h\*d\)/\(float\)\(Max_stage2\);", r"^\tint factor_t=1;", r"^\s*$", [],
     "namespace ref_host { using std::log2; using std::sqrt; using std::ceil;\n"
     "static int ref_schedule(int w, int h, int d, int start_res, int Max_stage2, int *whd)\n{\n"
     "    int n = 0; whd[0] = w; whd[1] = h; whd[2] = d; n++;        // append_new(w,h,d): level 0 (pyramid.cu:219)\n",
     "    int factor_t = 1;\n    for (int el = 0; el < maxl; el++)\n    {\n        whd[3 * n] = w; whd[3 * n + 1] = h; whd[3 * n + 2] = d; n++;   // append_new(w,h,d) (pyramid.cu:239)\n"),
    ("Algorithm/pyramid.cu", (r"\(float\)\(Max_stage2\);", r"^\t\tif\(maxl-el<=el_x\) w=ceil"), r"^\tfor\(int i=m_data\.size\(\)-2", r"^\s*$|^\t\}\s*$", [],
     "", "    }\n    (void)factor_t;\n    return n;\n}\n}\n"),
    # isignbit, calc_border, ssim                                   morph.cu:35-118
    ("Algorithm/morph.cu", r"^__device__ int isignbit\(", r"^// Level processing", None),
    # INIT_* constants, kernel_initialize_level, init_improving_mask  morph.cu:170-261
    ("Algorithm/morph.cu", r"^const int INIT_BW", r"^void Morph::initialize_level", None),
    # OPT_* constants ... kernel_optimize_level                     morph.cu:594-1345 (stops before the host helper addressof)
    ("Algorithm/morph.cu", r"^const int OPT_BW", r"^T \*addressof\(", r"^template <class T>\s*$|^\s*$"),
    # temp_ref, interpolate_temp_ref, smooth, fill_zeros_x/y, kernel_initialize_temp   upsample.cu:28-211
    ("Algorithm/upsample.cu", r"^__global__ void temp_ref\(", r"^void initialize_temp\(", None),
    # kernel_render_halfway_image                                   render.cu:16-60
    ("Algorithm/render.cu", r"^__global__ void kernel_render_halfway_image\(", r"^void render_halfway_image\(", None),
]


def cut(text, first, stop, drop, fname, subs=(), prefix="", suffix=""):
    lines = text.split("\n")
    start = 0
    if isinstance(first, tuple):                      # (anchor to search after, first line taken)
        after, first = first
        start = next((i for i, l in enumerate(lines) if re.search(after, l)), None)
        if start is None:
            raise SystemExit(f"{fname}: anchor {after!r} not found")
    a = next((i for i in range(start, len(lines)) if re.search(first, lines[i])), None)
    if a is None:
        raise SystemExit(f"{fname}: anchor {first!r} not found")
    b = next((i for i in range(a + 1, len(lines)) if re.search(stop, lines[i])), None)
    if b is None:
        raise SystemExit(f"{fname}: stop anchor {stop!r} not found")
    while drop and b > a and re.search(drop, lines[b - 1]):
        b -= 1
    body = lines[a:b]
    for pat, rep in subs:
        hits = [i for i, l in enumerate(body) if re.search(pat, l)]
        if len(hits) != 1:
            raise SystemExit(f"{fname}: substitution anchor {pat!r} matched {len(hits)} lines")
        body[hits[0]] = re.sub(pat, rep, body[hits[0]])
    return f"// ---- {fname}:{a + 1}-{b} (extracted at build time) ----\n" + prefix + "\n".join(body) + "\n" + suffix


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("REF", "/root/reference"))
    ap.add_argument("--keep", action="store_true", help="keep the temporary translation unit (prints its path)")
    args = ap.parse_args()
    ref = args.ref
    if not os.path.isdir(os.path.join(ref, "Algorithm")):
        print(f"reference tree absent ({ref}): keeping prebuilt oracle/_ref/libref_devfn.so")
        return 0
    os.makedirs(OUT_DIR, exist_ok=True)
    parts = ['#include <cmath>\n#include <ctime>\n#include "simt.h"\n#include <util/dmath.h>\n#include <util/linalg.h>\n#include "stencils.h"\n#include "Pyramid.h"\n#include <util/dimage.h>\n#include <util/box_sampler.h>\n#include "prelude.h"\n']
    for span in SPANS:
        fname, first, stop, drop = span[:4]
        with open(os.path.join(ref, fname), encoding="latin-1") as f:
            parts.append(cut(f.read().replace("\r\n", "\n"), first, stop, drop, fname, span[4] if len(span) > 4 else (),
                             span[5] if len(span) > 5 else "", span[6] if len(span) > 6 else ""))
    parts.append('#include "capi.inc"\n')
    tmp = tempfile.mkdtemp(prefix="refdev_")
    tu = os.path.join(tmp, "refdev_tu.cpp")
    with open(tu, "w") as f:
        f.write("\n".join(parts))
    pre = os.path.join(tmp, "pre.h")
    with open(pre, "w") as f:
        # linalg.hpp:369 names an undeclared `mat` inside a template body MSVC never checks; __min / __max are MSVC macros
        f.write("#include <cstring>\nstatic int mat;\n#define __min(a,b) ((a)<(b)?(a):(b))\n#define __max(a,b) ((a)>(b)?(a):(b))\n")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    inc = ["-I" + HERE, "-I" + os.path.join(HERE, "stubs"), "-I" + os.path.join(ref, "include"), "-I" + os.path.join(ref, "Algorithm"),
           "-I/usr/local/cuda/include"]
    # IEEE fp32, no contraction: the arithmetic contract of DESIGN.md section 2 (the reference binary itself was built with
    # -use_fast_math, which no CPU can reproduce)
    flags = ["-O2", "-std=c++17", "-fPIC", "-w", "-DNDEBUG", "-DCUDA_SM=35", "-ffp-contract=off", "-fno-fast-math", "-include", pre]
    # -Bsymbolic: the host stand-ins for cudaMemcpy / cudaMemset defined in capi.inc must be the ones this library calls, also
    # in a process that has the real CUDA runtime loaded (torch)
    cmd = [cxx] + flags + inc + ["-shared", "-Wl,-Bsymbolic", "-o", OUT, tu, os.path.join(ref, "Algorithm", "stencils.cpp")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-6000:])
        print("temporary translation unit kept at", tu)
        raise SystemExit("g++ failed building libref_devfn.so")
    if args.keep:
        print("translation unit:", tu)
    else:
        os.remove(tu); os.remove(pre); os.rmdir(tmp)
    print(os.path.normpath(OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())
