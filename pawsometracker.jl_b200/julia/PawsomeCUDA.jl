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
#   trckr.sz, trckr.radii, trckr.img.data                     # img.data is what read!(vid, …) fills (:166)
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

# `trckr.img.data` must stay a writable W×H Matrix{UInt8}-compatible buffer
# (the reference's frame is a PermutedDimsArray over a W×H Matrix{Gray{N0f8}},
# i.e. row-major H×W bytes — src/PawsomeTracker.jl:36).
struct FrameView{M<:AbstractMatrix}
    data::M
end

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
        data = parent(_img)                               # W×H matrix whose memory is row-major H×W
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pt_tracker_create, LIB), Cint,
                    (Cint, Cint, Cdouble, Cint, Cint, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}),
                    sz[1], sz[2], target_width, window_size[1], window_size[2],
                    darker_target ? 1 : 0, PT_PIX_U8, 0, h))
        t = new(sz, radii, FrameView(data), h[], 0)
        finalizer(t) do x
            x.handle != C_NULL && ccall((:pt_tracker_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle)
            x.handle = C_NULL
        end
        # fillvalue = mode(_img) of THIS frame (:47): upload once, histogram on the device
        GC.@preserve data begin
            check(ccall((:pt_tracker_set_frame, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
                        t.handle, pointer(data), sz[2]))
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
    data = trckr.img.data
    GC.@preserve data begin
        check(ccall((:pt_tracker_step_host, LIB), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{Cfloat}),
                    trckr.handle, pointer(data), trckr.sz[2], guess[1], guess[2], oi, oj, resp))
    end
    return (Int(oi[]), Int(oj[]))
end

end # module
