// pybind11 module `torch_bnb_fp4_ext` for B200: the reference's C++-extension op surface
// (reference csrc/torch_fp4.cpp:125-139 - same module name, same function names, same positional signatures,
// same ScalarType enum with exported values) implemented over the C-ABI of libfp4_b200.so (include/fp4_b200.h).
// The reference's own Python module (torch_bnb_fp4/__init__.py:11-18) imports this name and nothing else from
// native code, so building this file is all it takes to run the unmodified reference module on these kernels.
//
// Deliberate differences from the reference binding: launches go to the CURRENT stream of
// the input's device under a device guard (the reference: legacy stream, no guard), errors raise, gemv_fp4
// honours the `datatype` codebook and fills every row of a batch <= 8, qlinear_codebook* dequantise the whole
// weight (the reference passes the byte count as the element count: csrc/torch_fp4.cpp:90,101).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <map>
#include <mutex>
#include <tuple>

#include "fp4_b200.h"

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")

enum class ScalarTypeEnum {  // order = the C-ABI dtype codes = the reference's enum order (csrc/torch_fp4.cpp:22-26)
    float16,
    float32,
    bfloat16,
};

static torch::ScalarType get_scalar_type(ScalarTypeEnum t) {  // csrc/torch_fp4.cpp:28-39
    switch (t) {
        case ScalarTypeEnum::float16: return torch::kFloat16;
        case ScalarTypeEnum::float32: return torch::kFloat32;
        case ScalarTypeEnum::bfloat16: return torch::kBFloat16;
        default: throw py::type_error("Unsupported scalar type");
    }
}
static int code_of(torch::ScalarType t) {
    if (t == torch::kFloat16) return FP4_B200_F16;
    if (t == torch::kFloat32) return FP4_B200_F32;
    if (t == torch::kBFloat16) return FP4_B200_BF16;
    throw py::type_error("Unsupported scalar type");
}
static void check(int status, const char* what) {
    TORCH_CHECK(status == 0, what, ": ", fp4_b200_status_string(status), " (status ", status, ")");
}
static void* cur_stream(const torch::Tensor& t) {
    return (void*)at::cuda::getCurrentCUDAStream(t.device().index()).stream();
}

// The integer tensor-core GEMV is only valid for the bitsandbytes FP4 table: whether a `datatype` tensor holds it
// is read back once per (storage, version) - outside CUDA-graph capture - and remembered.
static bool code_is_bnb_fp4(const torch::Tensor& code) {
    static std::mutex mu;
    static std::map<std::tuple<const void*, int64_t>, bool> cache;
    if (code.numel() != 16 || code.scalar_type() != torch::kFloat32) return false;
    const auto key = std::make_tuple((const void*)code.data_ptr(), (int64_t)code._version());
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing((cudaStream_t)cur_stream(code), &cap);
    if (cap != cudaStreamCaptureStatusNone) return false;  // cannot read the table now: the generic kernel honours it
    static const float ref[16] = {0.0f, 5.208333333e-03f, 0.66666667f, 1.0f, 0.33333333f, 0.5f, 0.16666667f, 0.25f,
                                  -0.0f, -5.208333333e-03f, -0.66666667f, -1.0f, -0.33333333f, -0.5f, -0.16666667f, -0.25f};
    const torch::Tensor h = code.detach().reshape({-1}).cpu();
    bool ok = true;
    for (int i = 0; i < 16; ++i) ok = ok && h.data_ptr<float>()[i] == ref[i];  // by value: +0.0 == -0.0
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = ok;
    return ok;
}

// scratch of the stream-K GEMV kernels: one zero-filled buffer per (device, stream), created outside capture
static torch::Tensor& workspace(const torch::Tensor& like, void* stream, size_t need) {
    static std::mutex mu;
    static std::map<std::tuple<int, void*>, torch::Tensor> ws;
    std::lock_guard<std::mutex> lock(mu);
    torch::Tensor& t = ws[std::make_tuple((int)like.device().index(), stream)];
    if (!t.defined() || (size_t)t.numel() < need) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing((cudaStream_t)stream, &cap);
        TORCH_CHECK(cap == cudaStreamCaptureStatusNone,
                    "GEMV workspace must be created before CUDA-graph capture: run the layer once eagerly first");
        t = torch::zeros({(int64_t)std::max<size_t>(need, 1 << 20)},
                         torch::TensorOptions().dtype(torch::kUInt8).device(like.device()));
    }
    return t;
}

// ---- reference csrc/torch_fp4.cpp:41-50 ----------------------------------------------------------------------
static torch::Tensor dequantize_fp4(torch::Tensor A, torch::Tensor absmax, int blocksize, int M, int N,
                                    ScalarTypeEnum o_type) {
    CHECK_CUDA(A); CHECK_CUDA(absmax); CHECK_CONTIGUOUS(A); CHECK_CONTIGUOUS(absmax);
    TORCH_CHECK(A.scalar_type() == torch::kUInt8, "A must be uint8");
    TORCH_CHECK(absmax.scalar_type() == torch::kFloat32, "absmax must be float32");
    const c10::cuda::CUDAGuard guard(A.device());
    const auto dt = get_scalar_type(o_type);
    torch::Tensor out = torch::empty({M, N}, torch::TensorOptions().dtype(dt).device(A.device()));
    check(fp4_b200_dequantize(A.data_ptr<uint8_t>(), absmax.data_ptr<float>(), nullptr, out.data_ptr(),
                              (int64_t)M * N, blocksize, code_of(dt), cur_stream(A)), "dequantize_fp4");
    return out;
}

// ---- reference csrc/torch_fp4.cpp:52-62 (n = number of ELEMENTS, torch_bnb_fp4/__init__.py:452) ---------------
static torch::Tensor dequantize_codebook_impl(const torch::Tensor& A, const torch::Tensor& absmax,
                                              const torch::Tensor& codebook, int M, int N, int blocksize, int64_t n,
                                              torch::ScalarType dt) {
    CHECK_CUDA(A); CHECK_CUDA(absmax); CHECK_CUDA(codebook);
    CHECK_CONTIGUOUS(A); CHECK_CONTIGUOUS(absmax); CHECK_CONTIGUOUS(codebook);
    TORCH_CHECK(A.scalar_type() == torch::kUInt8, "A must be uint8");
    TORCH_CHECK(absmax.scalar_type() == torch::kFloat32 && codebook.scalar_type() == torch::kFloat32,
                "absmax and codebook must be float32");
    TORCH_CHECK(codebook.numel() == 16, "codebook must have 16 entries");
    TORCH_CHECK(n >= 0 && n <= (int64_t)M * N, "n out of range");
    const c10::cuda::CUDAGuard guard(A.device());
    torch::Tensor out = torch::empty({M, N}, torch::TensorOptions().dtype(dt).device(A.device()));
    check(fp4_b200_dequantize(A.data_ptr<uint8_t>(), absmax.data_ptr<float>(), codebook.data_ptr<float>(),
                              out.data_ptr(), n, blocksize, code_of(dt), cur_stream(A)), "dequantize_fp4_codebook");
    return out;
}
static torch::Tensor dequantize_fp4_codebook(torch::Tensor A, torch::Tensor absmax, torch::Tensor codebook, int M,
                                             int N, int blocksize, int n, ScalarTypeEnum dtype) {
    return dequantize_codebook_impl(A, absmax, codebook, M, N, blocksize, n, get_scalar_type(dtype));
}

// ---- reference csrc/torch_fp4.cpp:64-103: dequant + linear (the whole weight) ----------------------
static torch::Tensor dequant_tree(const torch::Tensor& A, const torch::Tensor& absmax, int M, int N, int blocksize,
                                  torch::ScalarType dt) {
    CHECK_CUDA(A); CHECK_CUDA(absmax); CHECK_CONTIGUOUS(A); CHECK_CONTIGUOUS(absmax);
    const c10::cuda::CUDAGuard guard(A.device());
    torch::Tensor out = torch::empty({M, N}, torch::TensorOptions().dtype(dt).device(A.device()));
    check(fp4_b200_dequantize(A.data_ptr<uint8_t>(), absmax.data_ptr<float>(), nullptr, out.data_ptr(),
                              (int64_t)M * N, blocksize, code_of(dt), cur_stream(A)), "qlinear");
    return out;
}
static torch::Tensor qlinear_(torch::Tensor A_in, torch::Tensor A, torch::Tensor absmax, int M, int N, int blocksize) {
    CHECK_CUDA(A_in);
    return torch::nn::functional::linear(A_in, dequant_tree(A, absmax, M, N, blocksize, A_in.scalar_type()));
}
static torch::Tensor qlinear_bias(torch::Tensor A_in, torch::Tensor A, torch::Tensor absmax, int M, int N,
                                  int blocksize, torch::Tensor bias) {
    CHECK_CUDA(A_in);
    return torch::nn::functional::linear(A_in, dequant_tree(A, absmax, M, N, blocksize, A_in.scalar_type()), bias);
}
static torch::Tensor qlinear_codebook(torch::Tensor A_in, torch::Tensor A, torch::Tensor absmax, torch::Tensor codebook,
                                      int M, int N, int blocksize) {
    CHECK_CUDA(A_in);
    return torch::nn::functional::linear(
        A_in, dequantize_codebook_impl(A, absmax, codebook, M, N, blocksize, (int64_t)M * N, A_in.scalar_type()));
}
static torch::Tensor qlinear_codebook_bias(torch::Tensor A_in, torch::Tensor A, torch::Tensor absmax,
                                           torch::Tensor codebook, int M, int N, int blocksize, torch::Tensor bias) {
    CHECK_CUDA(A_in);
    return torch::nn::functional::linear(
        A_in, dequantize_codebook_impl(A, absmax, codebook, M, N, blocksize, (int64_t)M * N, A_in.scalar_type()), bias);
}

// ---- reference csrc/torch_fp4.cpp:105-123 -> csrc/gemv_fp4_optimized.cu:277-368 -------------------------------
static torch::Tensor gemv_impl(const torch::Tensor& A, const torch::Tensor& B, const torch::Tensor& absmax,
                               const torch::Tensor& datatype, int blocksize, torch::ScalarType dt,
                               const std::vector<uint32_t>& Bshape, const c10::optional<torch::Tensor>& bias) {
    CHECK_CUDA(A); CHECK_CUDA(B); CHECK_CUDA(absmax); CHECK_CUDA(datatype);
    CHECK_CONTIGUOUS(A); CHECK_CONTIGUOUS(B); CHECK_CONTIGUOUS(absmax); CHECK_CONTIGUOUS(datatype);
    TORCH_CHECK(Bshape.size() == 2, "Bshape must be [out_features, in_features]");
    TORCH_CHECK(A.scalar_type() == dt, "A's dtype does not match the dtype argument");
    TORCH_CHECK(B.scalar_type() == torch::kUInt8, "B must be uint8");
    TORCH_CHECK(absmax.scalar_type() == torch::kFloat32 && datatype.scalar_type() == torch::kFloat32,
                "absmax and datatype must be float32");
    const int64_t n_out = Bshape[0], k = Bshape[1];
    TORCH_CHECK(A.dim() >= 1 && A.size(-1) == k, "A must be [..., in_features]");
    TORCH_CHECK(B.numel() * 2 >= n_out * k, "B holds fewer than N*K/2 bytes");
    const int64_t batch = k ? A.numel() / k : 0;
    auto sizes = A.sizes().vec();
    sizes.back() = n_out;
    const c10::cuda::CUDAGuard guard(A.device());
    torch::Tensor out = torch::empty(sizes, A.options());
    if (batch == 0 || n_out == 0) return out;
    const void* pbias = nullptr;
    if (bias.has_value() && bias->defined()) {
        CHECK_CUDA((*bias)); CHECK_CONTIGUOUS((*bias));
        TORCH_CHECK(bias->scalar_type() == dt && bias->numel() == n_out, "bias must be [out_features] of A's dtype");
        pbias = bias->data_ptr();
    }
    void* st = cur_stream(A);
    const unsigned flags = code_is_bnb_fp4(datatype) ? FP4_B200_FLAG_CODE_IS_BNB_FP4 : 0u;
    torch::Tensor& ws = workspace(A, st, fp4_b200_gemv_workspace_bytes((int)n_out));
    check(fp4_b200_gemv(A.data_ptr(), B.data_ptr<uint8_t>(), absmax.data_ptr<float>(), nullptr,
                        datatype.data_ptr<float>(), pbias, out.data_ptr(), (int)batch, (int)n_out, (int)k, blocksize,
                        code_of(dt), flags, ws.data_ptr(), (size_t)ws.numel(), st), "gemv_fp4");
    return out;
}
static torch::Tensor gemv_fp4(torch::Tensor A, torch::Tensor B, torch::Tensor absmax, torch::Tensor datatype,
                              int blocksize, ScalarTypeEnum dtype, std::vector<uint32_t> Bshape) {
    return gemv_impl(A, B, absmax, datatype, blocksize, get_scalar_type(dtype), Bshape, c10::nullopt);
}
// extension: bias fused into the epilogue (replaces the separate `out += bias`, torch_bnb_fp4/__init__.py:608-613)
static torch::Tensor gemv_fp4_bias(torch::Tensor A, torch::Tensor B, torch::Tensor absmax, torch::Tensor datatype,
                                   int blocksize, ScalarTypeEnum dtype, std::vector<uint32_t> Bshape,
                                   c10::optional<torch::Tensor> bias) {
    return gemv_impl(A, B, absmax, datatype, blocksize, get_scalar_type(dtype), Bshape, bias);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    pybind11::enum_<ScalarTypeEnum>(m, "ScalarType")
        .value("bfloat16", ScalarTypeEnum::bfloat16)
        .value("float16", ScalarTypeEnum::float16)
        .value("float32", ScalarTypeEnum::float32)
        .export_values();
    m.def("dequantize_fp4", &dequantize_fp4, "blockwise FP4 dequant, bitsandbytes constants");
    m.def("dequantize_fp4_codebook", &dequantize_fp4_codebook, "blockwise FP4 dequant through a 16-entry codebook");
    m.def("gemv_fp4", &gemv_fp4, "fused dequant + GEMV, batch 1..8");
    m.def("gemv_fp4_bias", &gemv_fp4_bias, "fused dequant + GEMV with the bias in the epilogue");
    m.def("qlinear", &qlinear_, "dequant + linear");
    m.def("qlinear_bias", &qlinear_bias, "dequant + linear with bias");
    m.def("qlinear_codebook", &qlinear_codebook, "codebook dequant + linear");
    m.def("qlinear_codebook_bias", &qlinear_codebook_bias, "codebook dequant + linear with bias");
    m.attr("backend") = "libfp4_b200 (sm_100a)";
}
