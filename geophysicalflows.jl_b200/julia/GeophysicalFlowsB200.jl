# GeophysicalFlowsB200.jl -- reference-side binding for libgfb (include/gfb.h).
#
# SOURCE ONLY: there is no Julia toolchain in the build environment, so this file is the stub a
# GeophysicalFlows.jl / FourierFlows.jl maintainer would add; the identical C ABI is exercised from
# Python (ctypes) by the test-suite.  It keeps the FourierFlows `Problem` / `stepforward!` / `calcN!` /
# `updatevars!` surface: a `B200TimeStepper` replaces the FourierFlows stepper object, and the
# methods below dispatch on it, so `stepforward!(prob, nsteps)` and the module functions
# (`energy(prob)`, ...) work unchanged on a `prob` built with `dev = GPU()`.
#
# Data crosses the boundary as raw device pointers of the CuArrays the Problem already owns:
#   sol::CuArray{ComplexF64}(nkr, nl[, nlayers])  <->  double2 sol[layer][l][kr]   (same bytes)
module GeophysicalFlowsB200

using CUDA, FourierFlows
import FourierFlows: stepforward!, dealias!

const libgfb = get(ENV, "LIBGFB", "libgfb.so")

const KIND_SINGLE, KIND_LAYERED, KIND_QL = Cint(0), Cint(1), Cint(2)   # KIND_QL: BarotropicQGQL
const STEPPERS = Dict("ForwardEuler" => 0, "FilteredForwardEuler" => 1, "RK4" => 2, "FilteredRK4" => 3,
                      "ETDRK4" => 4, "FilteredETDRK4" => 5, "AB3" => 6, "FilteredAB3" => 7, "LSRK54" => 8, "FilteredLSRK54" => 9)

# mirrors `struct gfb_config` field for field (include/gfb.h)
Base.@kwdef mutable struct GfbConfig
  struct_size::Cint = 0   # filled by `create` with sizeof(GfbConfig); gfb_create rejects a mismatch
  kind::Cint = 0;  nx::Cint = 0;  ny::Cint = 0;  nlayers::Cint = 1
  Lx::Cdouble = 2π; Ly::Cdouble = 2π
  kr_alias0::Cint = 0; l_alias0::Cint = 0; l_alias1::Cint = 0
  stepper::Cint = 2; dt::Cdouble = 0.0
  L_is_complex::Cint = 0
  L::Ptr{Cvoid} = C_NULL; filter::Ptr{Cdouble} = C_NULL
  expLdt::Ptr{Cvoid} = C_NULL; exphalfLdt::Ptr{Cvoid} = C_NULL; zeta::Ptr{Cvoid} = C_NULL
  alpha::Ptr{Cvoid} = C_NULL; beta::Ptr{Cvoid} = C_NULL; Gamma::Ptr{Cvoid} = C_NULL
  G::Ptr{Cdouble} = C_NULL; eta::Ptr{Cdouble} = C_NULL; Nconst::Ptr{Cvoid} = C_NULL
  Sinv::Ptr{Cdouble} = C_NULL; S::Ptr{Cdouble} = C_NULL; U::Ptr{Cdouble} = C_NULL
  Qx::Ptr{Cdouble} = C_NULL; Qy::Ptr{Cdouble} = C_NULL
  mu::Cdouble = 0.0; no_dealias_in_calcN::Cint = 0
  rank::Cint = 0; nranks::Cint = 0   # single GPU; a sharded driver sets both (include/gfb.h)
end

# the mirror above must be byte-identical to the C struct: checked once against the library itself
function check_config_layout()
  n = ccall((:gfb_config_size, libgfb), Cint, ())
  n == sizeof(GfbConfig) || error("libgfb: struct gfb_config is $n bytes in the library, $(sizeof(GfbConfig)) in GfbConfig")
end

mutable struct B200TimeStepper{T} <: FourierFlows.AbstractTimeStepper{T}
  ctx    :: Ptr{Cvoid}
  name   :: String
  filter :: Union{Nothing, Array{Float64}}   # kept so `prob.timestepper.filter` still works
end

check(rc, ctx=C_NULL) = rc == 0 ? nothing :
  error("libgfb: ", unsafe_string(ccall((:gfb_last_error, libgfb), Cstring, (Ptr{Cvoid},), ctx)))

"""
    B200TimeStepper(stepper::String, equation, dt, grid, tables...)

Built by each module's `Problem(dev::GPU; ...)` in place of `FourierFlows.TimeStepper(...)`: the host-side
tables the reference already computes (`equation.L`, `makefilter(equation)`, `getetdcoeffs(dt, L)`,
`grid.invKrsq` / `params.ψhfrombh` / `params.S⁻¹`, `params.Qx`, ...) are handed over once, as HOST arrays.
"""
function B200TimeStepper(stepper::String, eqn, dt, grid; kind, G=nothing, eta=nothing, Nconst=nothing,
                         Sinv=nothing, S=nothing, U=nothing, Qx=nothing, Qy=nothing, μ=0.0)
  L = Array(eqn.L)
  filtered = startswith(stepper, "Filtered")
  filter = filtered ? Array(FourierFlows.makefilter(eqn))[:, :, 1] : nothing
  check_config_layout()
  cfg = GfbConfig(struct_size=sizeof(GfbConfig), kind=kind, nx=grid.nx, ny=grid.ny, nlayers=size(L, 3), Lx=grid.Lx, Ly=grid.Ly,
                  kr_alias0=first(grid.kralias) - 1, l_alias0=first(grid.lalias) - 1, l_alias1=last(grid.lalias),
                  stepper=STEPPERS[stepper], dt=dt, L_is_complex=eltype(L) <: Complex, mu=μ)
  keep = Any[L]
  hp(a) = a === nothing ? C_NULL : (b = Array(a); push!(keep, b); Ptr{Cvoid}(pointer(b)))
  cfg.L = hp(L); cfg.filter = hp(filter)
  if occursin("ETDRK4", stepper)
    ζ, α, β, Γ = FourierFlows.getetdcoeffs(dt, L)
    cfg.expLdt = hp(exp.(dt .* L)); cfg.exphalfLdt = hp(exp.(dt .* L ./ 2))
    cfg.zeta = hp(ζ); cfg.alpha = hp(α); cfg.beta = hp(β); cfg.Gamma = hp(Γ)
  end
  cfg.G = hp(G); cfg.eta = hp(eta); cfg.Nconst = hp(Nconst); cfg.U = hp(U); cfg.Qx = hp(Qx); cfg.Qy = hp(Qy)
  # params.S⁻¹ is Array{SMatrix{n,n}}(nkr, nl), column-major inside each SMatrix: the ABI wants row-major
  rowmajor(M) = M === nothing ? nothing : [M[m, l][i, j] for j in 1:size(M[1,1], 2), i in 1:size(M[1,1], 1), m in axes(M, 1), l in axes(M, 2)]
  cfg.Sinv = hp(rowmajor(Sinv)); cfg.S = hp(rowmajor(S))
  ctx = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve keep check(ccall((:gfb_create, libgfb), Cint, (Ref{Ptr{Cvoid}}, Ref{GfbConfig}), ctx, cfg))
  ts = B200TimeStepper{eltype(L)}(ctx[], stepper, filter)
  finalizer(t -> ccall((:gfb_destroy, libgfb), Cint, (Ptr{Cvoid},), t.ctx), ts)
  return ts
end

stream() = Ptr{Cvoid}(CUDA.stream().handle)
dptr(a::CuArray) = Ptr{Cvoid}(UInt(pointer(a)))
dptr(::Nothing) = C_NULL

# stepforward!(prob, nsteps): FourierFlows' loop `for _ in 1:nsteps; stepforward!(sol, clock, ts, ...); end`
# becomes ONE call when the forcing is absent, frozen (stochastic) or time independent.
function stepforward!(prob::FourierFlows.Problem{<:Any, <:Any, <:B200TimeStepper}, nsteps::Int=1)
  ts, v, c = prob.timestepper, prob.vars, prob.clock
  Fh      = hasproperty(v, :Fh) ? v.Fh : (hasproperty(v, :Fqh) ? v.Fqh : nothing)
  prevsol = hasproperty(v, :prevsol) ? v.prevsol : nothing
  if Fh === nothing
    check(ccall((:gfb_step, libgfb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                ts.ctx, dptr(prob.sol), nsteps, C_NULL, C_NULL, stream()), ts.ctx)
    for _ in 1:nsteps; c.t += c.dt; c.step += 1; end
  elseif prevsol !== nothing                       # stochastic: F drawn once per step at t == clock.t
    calcF! = hasproperty(prob.params, :calcF!) ? prob.params.calcF! : prob.params.calcFq!
    for i in 1:nsteps
      dealias!(prob.sol, prob.grid)
      calcF!(Fh, prob.sol, c.t, c, v, prob.params, prob.grid)
      # the aliased band of the returned state is only needed when somebody can look at it (last step of this call;
      # the diagnostics variant of stepforward! switches it on whenever a diagnostic is due)
      ccall((:gfb_set_defer_alias_band, libgfb), Cint, (Ptr{Cvoid}, Cint), ts.ctx, i < nsteps ? 1 : 0)
      check(ccall((:gfb_step, libgfb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                  ts.ctx, dptr(prob.sol), 1, dptr(Fh), dptr(prevsol), stream()), ts.ctx)
      c.t += c.dt; c.step += 1
    end
    ccall((:gfb_set_defer_alias_band, libgfb), Cint, (Ptr{Cvoid}, Cint), ts.ctx, 0)
  else                                             # deterministic forcing: re-evaluated at every stage time
    calcF! = hasproperty(prob.params, :calcF!) ? prob.params.calcF! : prob.params.calcFq!
    nst = ccall((:gfb_nstages, libgfb), Cint, (Ptr{Cvoid},), ts.ctx)
    cs = nst == 4 ? (0.0, 0.5, 0.5, 1.0) : (0.0,)
    for _ in 1:nsteps
      for i in 0:nst-1
        calcF!(Fh, prob.sol, c.t + cs[i+1] * c.dt, c, v, prob.params, prob.grid)
        check(ccall((:gfb_stage, libgfb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    ts.ctx, dptr(prob.sol), i, dptr(Fh), C_NULL, stream()), ts.ctx)
      end
      c.t += c.dt; c.step += 1
    end
  end
  return nothing
end

# the calcN! seam: `equation.calcN!(N, sol, t, clock, vars, params, grid)` (dealiases `sol` in place)
function calcN!(N, sol, t, clock, vars, params, grid, ts::B200TimeStepper; Fh=nothing)
  check(ccall((:gfb_calcN, libgfb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              ts.ctx, dptr(N), dptr(sol), dptr(Fh), stream()), ts.ctx)
end

# mul!(ŷ, grid.rfftplan, x) / ldiv!(x, grid.rfftplan, ŷ)
rfft!(out, x, ts::B200TimeStepper) =
  check(ccall((:gfb_rfft, libgfb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
              ts.ctx, dptr(out), dptr(x), size(x, 3), stream()), ts.ctx)
irfft!(out, xh, ts::B200TimeStepper) =
  check(ccall((:gfb_irfft, libgfb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
              ts.ctx, dptr(out), dptr(xh), size(xh, 3), stream()), ts.ctx)

# updatevars!(prob): dealias + inversion + the c2r transforms, no deepcopy
updatevars!(prob, ts::B200TimeStepper; q=nothing, ψ=nothing, u=nothing, v=nothing) =
  check(ccall((:gfb_updatevars, libgfb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              ts.ctx, dptr(prob.sol), dptr(q), dptr(ψ), dptr(u), dptr(v), stream()), ts.ctx)

# parsevalsum-type diagnostics: raw weighted sum from the library, prefactors as in the reference
function parsevalsum_b200(ts::B200TimeStepper, kind::Integer, sol, grid; aux=nothing, aux2=nothing, p=0.0, nout=1)
  out = zeros(Float64, 16)
  check(ccall((:gfb_diag, libgfb), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}, Cint, Ptr{Cvoid}),
              ts.ctx, kind, dptr(sol), dptr(aux), dptr(aux2), p, out, 16, stream()), ts.ctx)
  return out[1:nout] .* (grid.Lx * grid.Ly / (grid.nx^2 * grid.ny^2))
end
# e.g. TwoDNavierStokes.energy(prob) = parsevalsum_b200(ts, 1, prob.sol, grid)[1] * 0.5 / (Lx*Ly)   (GFB_DIAG_INVK_ABS2)

end # module
