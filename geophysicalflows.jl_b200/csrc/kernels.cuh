// kernels.cuh -- the three fused passes of one calcN! + stage update, and the plain 2-D transforms.
//
//   K1  spectral pre-pass (dealias mask on load, psi inversion / S^-1 matvec, i*k derivatives, 1/(nx*ny)) fused
//       with the inverse FFT along y (columns)
//         k_col_inv_ws       single-layer modules at ny = 4096 (the headline): warp-specialised, one CTA per SM = an FFT
//                            group + a helper group that reads whole sectors, prepares both columns of a unit ahead of the
//                            transforms and hands them over through TMEM
//         k_col_inv_single   single-layer modules, other sizes: persistent, rolling cp.async prefetch, TMEM stash
//         k_col_inv          MultiLayerQG pre-pass and plain inverse transforms
//         k_col_inv16k_dit   ny = 16384 (default): lane-paired columns (lane 2j + c = column c), 4 x 4096 decimated in time:
//                            every row read once, natural-order stores (512 contiguous bytes per warp: NVLink-friendly)
//         k_col_inv16k       ny = 16384, GFB_K1_16K_DIT=0: 4 x 4096 by a radix-4 DIF step while loading, four q units per pair
//   K2  Hermitian c2r along x of (u, v) and q, physical products, r2c along x -- physical fields never touch HBM
//         k_row_product_tma  nx = 4096 (the headline) and 8192, one GPU: as k_row_product with TMA row gathers / scatters
//         k_row_product_ws   nx = 4096, opt-in (GFB_ROW_WS=1; measured slower): warp-specialised -- the FFT group does the five
//                            transforms of a row pair; the helper group gathers rows by TMA, assembles the Hermitian input,
//                            separates the output spectra and scatters the result rows by TMA; hand-off through TMEM
//         k_row_product      other sizes / layered: persistent, line-buffer prefetch of the next phase, q parked in TMEM
//         k_row              plain c2r / r2c (ldiv! / mul! seams, updatevars!)
//         k_row_product_half / k_row_half   nx = 16384: every real row as one 8192-point half-length transform; all
//                            per-thread twiddles on chip (TMEM + shared-memory tables): one CTA per SM hides no L2 round trip
//   K3  forward FFT along y of the two product fields fused with the divergence (-i kr, -i l), forcing and the
//       time-stepper stage update
//         k_col_fwd          every register-resident size incl. the headline: one transform call site over the four columns,
//                            TMEM-parked partners, kind-specialised fused epilogue
//         k_col_fwd_ws       ny = 4096, opt-in (GFB_COL_WS3=1; measured slower): warp-specialised -- the FFT group transforms
//                            while the helper group streams the next columns in and runs the previous pair's epilogue
//         k_col_fwd16k (+ _plain)   ny = 16384; k_col_fwd16k_lp: lane-paired variant, opt-in (GFB_K3_16K_LP=1; measured slower)
//         k_col_inv_dft / k_col_fwd_dft (cols_dft.cu)   any even ny <= 1024 that is not a power of two
//
// Reference sequence being replaced: calcN_advection! /root/reference/src/twodnavierstokes.jl:253-275,
// src/surfaceqg.jl:257-278, src/singlelayerqg.jl:356-432, src/multilayerqg.jl:695-731; dealias!
// (FourierFlows) called at src/twodnavierstokes.jl:287; stage updates of FourierFlows' RK4/ETDRK4/...
// steppers.
//
// Layouts: spectral `cplx a[layer][l][kr]` (kr contiguous), physical
// `double a[layer][y][x]`.  Work arrays between the passes: W[field][y][kr] (ld = ldw) and
// P[field][y][kr] (ld = ldp), ctx-owned.
#pragma once
#include "layout.cuh"
#include "k1_col_inv.cuh"
#include "k1_col_inv_ws.cuh"
#include "k2_row.cuh"
#include "k2_row_ws.cuh"
#include "k2_row_half.cuh"
#include "k3_col_fwd.cuh"
#include "k3_col_fwd_ws.cuh"
#include "cols16k.cuh"
#include "aux_kernels.cuh"
