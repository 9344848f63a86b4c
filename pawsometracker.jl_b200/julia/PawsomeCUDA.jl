# PawsomeCUDA.jl — drop-in replacement of PawsomeTracker's `Tracker` (reference
# src/PawsomeTracker.jl:32-62) that routes the DoG-window filter + findmax step
# through libpawsome_cuda.so (C ABI in include/pawsome.h).
#
# NOT EXECUTED IN THIS BUILD ENVIRONMENT: the image has no `julia`.  The file is
# kept deliberately thin and mirrors, call for call, the Python ctypes harness
# (pawsometracker.jl_b200/tracker.py) that the GPU parity tests drive.  See
# INTEGRATION.md for the three-line patch to src/PawsomeTracker.jl.
#
# Surface preserved for the rest of the package:
#   Tracker(img, target_width, window_size::NTuple{2,Int}, darker_target::Bool)
#   (trckr::Tracker)(guess::NTuple{2,Int})::NTuple{2,Int}     # 1-based (row, col), clamped
#   trckr.sz, trckr.radii, trckr.img.data                     # img.data: the H×W frame view read!(vid, …) fills (:166)
#                                                             # and dia(trckr.img.data, ij) resizes (:168)
module PawsomeCUDA

export Tracker

const LIB = get(ENV, "PAWSOME_CUDA_LIB", "libpawsome_cuda.so")

const PT_PIX_U8 = Cint(0)

struct PawsomeCUDAError <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::PawsomeCUDAError) = print(io, "libpawsome_cuda error ", e.code, ": ", e.msg)

last_error() = unsafe_string(ccall((:pt_last_error, LIB), Cstring, ()))

function check(code::Cint)
    code < 0 && throw(code == -1 ? ArgumentError(last_error()) : PawsomeCUDAError(code, last_error()))
    return code
end

# `trckr.img.data` keeps the reference's meaning: the H×W view of the frame (`_img`, a
# `PermutedDimsArray{Gray{N0f8},2,(2,1)}` over a W×H `Matrix`, src/PawsomeTracker.jl:36) — the array that
# `read!(vid, trckr.img.data)` fills (:166) and that `dia(trckr.img.data, ij)` hands to `imresize!`
# (:168, src/diagnose.jl:33).  In the reference `trckr.img` is the PaddedView and `.data` its parent; here the
# padding happens inside the kernels, so `img` is a one-field wrapper with the same property.  Only the `ccall`
# sites reach for `parent(data)`: the W×H matrix whose memory IS the row-major H×W byte frame the library reads.
struct FrameView{A<:AbstractMatrix}
    data::A
end

# The memory behind an H×W frame as the library wants it: H rows of W bytes, row-major.  A PermutedDimsArray with
# permutation (2,1) over a dense W×H matrix is exactly that; anything else (a plain column-major H×W Matrix would
# silently be read transposed) is refused.
function rowmajor_parent(img::PermutedDimsArray{T, 2, (2, 1)}) where {T}
    p = parent(img)
    p isa DenseMatrix || throw(ArgumentError("the frame's parent must be a dense W×H matrix"))
    sizeof(T) == 1 || throw(ArgumentError("the frame's element type must be 1 byte (Gray{N0f8} / UInt8)"))
    return p
end
rowmajor_parent(img) = throw(ArgumentError(
    "Tracker needs the frame as VideoIO delivers it: a PermutedDimsArray{<:Any,2,(2,1)} over a W×H matrix " *
    "(row-major H×W bytes); got $(typeof(img))"))

mutable struct Tracker
    sz::Tuple{Int, Int}
    radii::Tuple{Int, Int}
    img::FrameView
    handle::Ptr{Cvoid}
    fillvalue::Int

    # src/PawsomeTracker.jl:39-52
    function Tracker(_img, target_width, window_size, darker_target)
        sz = size(_img)                                   # (H, W)
        radii = window_size .÷ 2
        mem = rowmajor_parent(_img)                       # W×H matrix = row-major H×W bytes
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pt_tracker_create, LIB), Cint,
                    (Cint, Cint, Cdouble, Cint, Cint, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
                    sz[1], sz[2], target_width, window_size[1], window_size[2],
                    darker_target ? 1 : 0, PT_PIX_U8, 0, h))
        t = new(sz, radii, FrameView(_img), h[], 0)       # img.data stays the H×W view (read!, dia)
        finalizer(t) do x
            x.handle != C_NULL && ccall((:pt_tracker_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle)
            x.handle = C_NULL
        end
        # fillvalue = mode(_img) of THIS frame (:47): upload once, histogram on the device
        GC.@preserve mem begin
            check(ccall((:pt_tracker_set_frame, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                        t.handle, pointer(mem), sz[2]))
        end
        fill = Ref{Cint}(0)
        check(ccall((:pt_tracker_compute_fill, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}), t.handle, fill))
        t.fillvalue = fill[]
        return t
    end
end

# src/PawsomeTracker.jl:55-62 — window = guess ± radii, DoG response, findmax, clamp.
# Only the window's footprint of the host frame crosses PCIe.
function (trckr::Tracker)(guess::NTuple{2, Int})
    oi = Ref{Cint}(0); oj = Ref{Cint}(0); resp = Ref{Cfloat}(0)
    mem = rowmajor_parent(trckr.img.data)                 # the bytes read!(vid, trckr.img.data) just wrote
    GC.@preserve mem begin
        check(ccall((:pt_tracker_step_host, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cfloat}),
                    trckr.handle, pointer(mem), trckr.sz[2], guess[1], guess[2], oi, oj, resp))
    end
    return (Int(oi[]), Int(oj[]))
end

end # module
