// torch_ext.cpp -- thin PyTorch shim over the C ABI (include/gstar_raster.h).
//
// Exposes the exact three functions of the reference's pybind module
// (DGR/ext.cpp:15-19; signatures DGR/rasterize_points.h:18-66) so that
// diff_gaussian_rasterization/__init__.py-style callers work unchanged:
//     rasterize_gaussians, rasterize_gaussians_backward, mark_visible
// PyTorch is used only for memory (caching allocator through the resize callbacks), the current
// stream and the device guard.  No computation happens here; there is no CPU fallback.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <tuple>

#include "../../include/gstar_raster.h"

namespace {

// mirrors resizeFunctional(), DGR/rasterize_points.cu:27-33
char* resize_cb(void* user, size_t nbytes)
{
    auto* t = reinterpret_cast<torch::Tensor*>(user);
    t->resize_({(long long)nbytes});
    return reinterpret_cast<char*>(t->data_ptr());
}

const float* fptr(const torch::Tensor& t) { return t.numel() == 0 ? nullptr : t.data_ptr<float>(); }

// [P,c] colours of a second pass -> contiguous [P,4] (one float4 per Gaussian; unused channels zero)
torch::Tensor pad4(const torch::Tensor& t)
{
    if (t.size(1) == 4) return t.contiguous();
    torch::Tensor out = torch::zeros({t.size(0), 4}, t.options());
    out.slice(1, 0, t.size(1)).copy_(t);
    return out;
}

torch::Tensor prep(const torch::Tensor& t, const torch::Device& dev)
{
    if (t.numel() == 0) return t;
    TORCH_CHECK(t.scalar_type() == torch::kFloat32, "rasterizer inputs must be float32");
    torch::Tensor r = t;
    if (r.device() != dev) r = r.to(dev);
    return r.contiguous();
}

void check(int rc)
{
    if (rc < 0) {
        if (rc == GSTAR_ERR_NONRGB) throw std::runtime_error(gstar_last_error());
        TORCH_CHECK(false, gstar_last_error());
    }
}

}  // namespace

// Set by the Python wrapper around a call none of whose inputs requires grad (inference): the forward then skips the hit
// log.  Thread-local, reset by the wrapper right after the call; the reference's 19-argument signature stays as it is.
static thread_local bool t_forward_only = false;
void set_forward_only(bool flag) { t_forward_only = flag; }

// second: (colors2 [P,3], background2 [3]) of a second feature pass blended in the same kernel, or nullptr; out2 receives its image
static std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> forward_impl(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors, const torch::Tensor& opacity,
    const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
    const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height,
    const int image_width, const torch::Tensor& sh, const int degree, const torch::Tensor& campos, const bool prefiltered,
    const bool debug, const std::pair<torch::Tensor, torch::Tensor>* second, torch::Tensor* out2)
{
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) {
        AT_ERROR("means3D must have dimensions (num_points, 3)");
    }
    TORCH_CHECK(means3D.is_cuda(), "gaustar_b200: means3D must be a CUDA tensor (there is no CPU path)");
    const torch::Device dev = means3D.device();
    c10::cuda::CUDAGuard guard(dev);
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(dev.index()).stream();

    const int P = means3D.size(0);
    const int H = image_height, W = image_width;
    auto float_opts = means3D.options().dtype(torch::kFloat32);
    auto byte_opts = torch::TensorOptions(torch::kByte).device(dev);

    torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
    torch::Tensor geomBuffer = torch::empty({0}, byte_opts);
    torch::Tensor binningBuffer = torch::empty({0}, byte_opts);
    torch::Tensor imgBuffer = torch::empty({0}, byte_opts);
    int rendered = 0;
    torch::Tensor out_color;
    if (P != 0) {
        out_color = torch::empty({3, H, W}, float_opts);
        int M = 0;
        if (sh.size(0) != 0) M = sh.size(1);
        const torch::Tensor bg = prep(background, dev), m3 = prep(means3D, dev), col = prep(colors, dev), op = prep(opacity, dev),
                            sc = prep(scales, dev), rot = prep(rotations, dev), cov = prep(cov3D_precomp, dev), vm = prep(viewmatrix, dev),
                            pm = prep(projmatrix, dev), shc = prep(sh, dev), cp = prep(campos, dev);
        gstar_fwd_args a = {};
        a.P = P; a.D = degree; a.M = M;
        a.background = fptr(bg); a.width = W; a.height = H;
        a.means3D = fptr(m3); a.shs = fptr(shc); a.colors_precomp = fptr(col); a.opacities = fptr(op);
        a.scales = fptr(sc); a.scale_modifier = scale_modifier; a.rotations = fptr(rot); a.cov3D_precomp = fptr(cov);
        a.viewmatrix = fptr(vm); a.projmatrix = fptr(pm); a.cam_pos = fptr(cp);
        a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy; a.prefiltered = prefiltered ? 1 : 0;
        a.out_color = out_color.data_ptr<float>(); a.radii = radii.data_ptr<int>(); a.debug = debug ? 1 : 0;
        a.forward_only = t_forward_only ? 1 : 0;
        torch::Tensor col2, bg2;
        if (second) {
            TORCH_CHECK(second->first.is_cuda() && second->first.dim() == 2 && second->first.size(0) == P && second->first.size(1) >= 1 &&
                            second->first.size(1) <= 4 && second->second.numel() == second->first.size(1),
                        "gaustar_b200: the second pass needs CUDA colours of shape (P, c), c = 1..4, and a background of c values");
            const int c2 = (int)second->first.size(1);
            col2 = pad4(prep(second->first, dev)); bg2 = prep(second->second, dev);
            *out2 = torch::empty({c2, H, W}, float_opts);
            a.colors2 = fptr(col2); a.background2 = fptr(bg2); a.out_color2 = out2->data_ptr<float>(); a.channels2 = c2;
        }
        rendered = gstar_raster_forward(&a, resize_cb, &geomBuffer, resize_cb, &binningBuffer, resize_cb, &imgBuffer, stream);
        if (rendered == GSTAR_ERR_NOLOG) return std::make_tuple(rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer);  // the caller re-blends instead
        check(rendered);
    } else {
        out_color = torch::zeros({3, H, W}, float_opts);  // rasterize_points.cu:66
        if (second) *out2 = torch::zeros({second->second.numel(), H, W}, float_opts);
    }
    return std::make_tuple(rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer);
}

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> RasterizeGaussiansCUDA(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors, const torch::Tensor& opacity,
    const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
    const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height,
    const int image_width, const torch::Tensor& sh, const int degree, const torch::Tensor& campos, const bool prefiltered,
    const bool debug)
{
    return forward_impl(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                        tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug, nullptr, nullptr);
}

// Two feature passes in one blend (gstar_fwd_args::colors2): the 19 arguments of rasterize_gaussians + the second pass's colours and
// background.  Returns (num_rendered, out_color, out_color2, radii, geomBuffer, binningBuffer, imgBuffer); num_rendered ==
// GSTAR_ERR_NOLOG (-5) means the view cannot have its hit log and nothing usable was rendered: re-blend the second pass instead.
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> RasterizeGaussiansDualCUDA(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors, const torch::Tensor& opacity,
    const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
    const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height,
    const int image_width, const torch::Tensor& sh, const int degree, const torch::Tensor& campos, const bool prefiltered,
    const bool debug, const torch::Tensor& colors2, const torch::Tensor& background2)
{
    const std::pair<torch::Tensor, torch::Tensor> second(colors2, background2);
    torch::Tensor out2;
    auto r = forward_impl(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                          tan_fovy, image_height, image_width, sh, degree, campos, prefiltered, debug, &second, &out2);
    return std::make_tuple(std::get<0>(r), std::get<1>(r), out2, std::get<2>(r), std::get<3>(r), std::get<4>(r), std::get<5>(r));
}

// Shared-geometry re-blend (gstar_raster_reblend): second pass over the Gaussians and camera of an earlier
// forward call with other colours.  Returns (num_rendered, out_color, binningBuffer, imgBuffer); the geometry buffer and
// radii of the source call serve the backward of this pass as well.
std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor> ReblendGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& colors,
                                                                                  const int image_height, const int image_width,
                                                                                  const torch::Tensor& srcBinningBuffer,
                                                                                  const torch::Tensor& srcImgBuffer, const bool debug,
                                                                                  const torch::Tensor& srcViewmatrix,
                                                                                  const torch::Tensor& srcProjmatrix,
                                                                                  const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix)
{
    TORCH_CHECK(colors.is_cuda() && colors.dim() == 2 && colors.size(1) == 3, "gaustar_b200: re-blend needs CUDA colors_precomp of shape (P, 3)");
    TORCH_CHECK(srcBinningBuffer.is_cuda() && srcImgBuffer.is_cuda() && srcImgBuffer.numel() > 0 && srcBinningBuffer.numel() > 0,
                "gaustar_b200: re-blend needs the source call's buffers");
    const torch::Device dev = colors.device();
    c10::cuda::CUDAGuard guard(dev);
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(dev.index()).stream();
    const int P = colors.size(0);
    auto byte_opts = torch::TensorOptions(torch::kByte).device(dev);
    torch::Tensor binningBuffer = torch::empty({0}, byte_opts);
    torch::Tensor imgBuffer = torch::empty({0}, byte_opts);
    torch::Tensor out_color = torch::empty({3, image_height, image_width}, colors.options().dtype(torch::kFloat32));
    const torch::Tensor bg = prep(background, dev), col = prep(colors, dev);
    gstar_reblend_args a;
    a.P = P; a.width = image_width; a.height = image_height;
    a.background = fptr(bg); a.colors_precomp = fptr(col);
    a.src_binning_buffer = reinterpret_cast<const char*>(srcBinningBuffer.data_ptr());
    a.src_image_buffer = reinterpret_cast<const char*>(srcImgBuffer.data_ptr());
    a.out_color = out_color.data_ptr<float>();
    a.debug = debug ? 1 : 0;
    a.forward_only = t_forward_only ? 1 : 0;
    // camera guard (checked on the device): all four 4x4 matrices or none (empty tensors)
    const bool cam_check = srcViewmatrix.numel() == 16 && srcProjmatrix.numel() == 16 && viewmatrix.numel() == 16 && projmatrix.numel() == 16;
    const torch::Tensor svm = cam_check ? prep(srcViewmatrix, dev) : srcViewmatrix, spm = cam_check ? prep(srcProjmatrix, dev) : srcProjmatrix,
                        cvm = cam_check ? prep(viewmatrix, dev) : viewmatrix, cpm = cam_check ? prep(projmatrix, dev) : projmatrix;
    a.src_viewmatrix = cam_check ? fptr(svm) : nullptr; a.src_projmatrix = cam_check ? fptr(spm) : nullptr;
    a.viewmatrix = cam_check ? fptr(cvm) : nullptr; a.projmatrix = cam_check ? fptr(cpm) : nullptr;
    const int rendered = gstar_raster_reblend(&a, resize_cb, &binningBuffer, resize_cb, &imgBuffer, stream);
    check(rendered);
    return std::make_tuple(rendered, out_color, binningBuffer, imgBuffer);
}

struct FusedTargets {
    torch::Tensor means3D, sh, opacity, scales, rotations;
    bool atomic = false;  // several backward calls (different streams) may be adding into these tensors at once
};
struct SecondPass {  // backward of a two-pass forward: the second image's upstream gradient, background, colours; [P] moment scratch of a 4th channel
    torch::Tensor dL_dout_color2, background2, colors2, scratch2;
};
using BwdTuple = std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>;
static BwdTuple backward_impl(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii, const torch::Tensor& colors,
                              const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier,
                              const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                              const float tan_fovx, const float tan_fovy, const torch::Tensor& dL_dout_color, const torch::Tensor& sh,
                              const int degree, const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                              const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const bool debug, const FusedTargets* fused,
                              const torch::Tensor* preloaded_scratch = nullptr, const SecondPass* second = nullptr);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                               const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                               const torch::Tensor& geomBuffer, const int R, const torch::Tensor& binningBuffer,
                               const torch::Tensor& imageBuffer, const bool debug)
{
    return backward_impl(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                         tan_fovy, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug, nullptr);
}

// Gradient-accumulation fusion (multi-view steps): the five parameter gradients are ADDED into the given tensors
// (typically the leaves' .grad = views of the flat all-reduce buffer) inside the per-Gaussian backward kernel instead of
// being written to fresh tensors and summed by autograd's AccumulateGrad afterwards (which re-reads and re-writes
// 236 MB per view at 1 M Gaussians / SH degree 3).  Returns the same 8-tuple; the five fused entries are the targets.
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardFusedCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                                    const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                                    const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                                    const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                                    const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                                    const torch::Tensor& geomBuffer, const int R, const torch::Tensor& binningBuffer,
                                    const torch::Tensor& imageBuffer, const bool debug, torch::Tensor acc_means3D, torch::Tensor acc_sh,
                                    torch::Tensor acc_opacity, torch::Tensor acc_scales, torch::Tensor acc_rotations, const bool atomic)
{
    const int P = means3D.size(0);
    const int M = sh.size(0) != 0 ? (int)sh.size(1) : 0;
    auto ok = [&](const torch::Tensor& t, int64_t n) {
        return t.is_cuda() && t.device() == means3D.device() && t.scalar_type() == torch::kFloat32 && t.is_contiguous() && t.numel() == n;
    };
    TORCH_CHECK(ok(acc_means3D, (int64_t)P * 3) && ok(acc_opacity, P) && ok(acc_scales, (int64_t)P * 3) && ok(acc_rotations, (int64_t)P * 4) &&
                    (M == 0 || ok(acc_sh, (int64_t)P * M * 3)),
                "gaustar_b200: fused accumulation targets must be contiguous float32 CUDA tensors of the parameters' sizes");
    FusedTargets ft{acc_means3D, acc_sh, acc_opacity, acc_scales, acc_rotations, atomic};
    return backward_impl(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                         tan_fovy, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug, &ft);
}

static std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
backward_impl(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii, const torch::Tensor& colors,
              const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
              const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
              const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
              const torch::Tensor& geomBuffer, const int R, const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
              const bool debug, const FusedTargets* fused, const torch::Tensor* preloaded_scratch, const SecondPass* second)
{
    TORCH_CHECK(means3D.is_cuda(), "gaustar_b200: means3D must be a CUDA tensor (there is no CPU path)");
    const torch::Device dev = means3D.device();
    c10::cuda::CUDAGuard guard(dev);
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(dev.index()).stream();
    const int P = means3D.size(0);
    const int H = dL_dout_color.size(1);
    const int W = dL_dout_color.size(2);
    int M = 0;
    if (sh.size(0) != 0) M = sh.size(1);
    auto opts = means3D.options().dtype(torch::kFloat32);
    // every output is fully written by the fused per-Gaussian backward kernel: no zero-fills
    // (the reference zero-fills nine tensors per call, rasterize_points.cu:150-158)
    torch::Tensor dL_dmeans3D = fused ? fused->means3D : torch::empty({P, 3}, opts);
    torch::Tensor dL_dmeans2D = torch::empty({P, 3}, opts);
    // gradients of inputs that were not given are intermediates of the fused kernel: they are not materialised
    // (the reference returns them too, rasterize_points.cu:190-195, but autograd drops gradients of absent inputs)
    const bool has_colors = colors.numel() != 0, has_cov = cov3D_precomp.numel() != 0;
    torch::Tensor dL_dcolors = torch::empty({has_colors ? P : 0, 3}, opts);
    torch::Tensor dL_dopacity = fused ? fused->opacity : torch::empty({P, 1}, opts);
    torch::Tensor dL_dcov3D = torch::empty({has_cov ? P : 0, 6}, opts);
    torch::Tensor dL_dsh = (fused && M) ? fused->sh : torch::empty({P, M, 3}, opts);
    torch::Tensor dL_dscales = fused ? fused->scales : torch::empty({P, 3}, opts);
    torch::Tensor dL_drotations = fused ? fused->rotations : torch::empty({P, 4}, opts);
    if (P != 0) {
        torch::Tensor scratch = preloaded_scratch ? *preloaded_scratch : torch::zeros({P, GSTAR_GRAD_SCRATCH_FLOATS}, opts);
        const torch::Tensor bg = prep(background, dev), m3 = prep(means3D, dev), col = prep(colors, dev), sc = prep(scales, dev),
                            rot = prep(rotations, dev), cov = prep(cov3D_precomp, dev), vm = prep(viewmatrix, dev),
                            pm = prep(projmatrix, dev), shc = prep(sh, dev), cp = prep(campos, dev), dpix = prep(dL_dout_color, dev);
        const torch::Tensor rad = radii.contiguous(), gb = geomBuffer.contiguous(), bb = binningBuffer.contiguous(),
                            ib = imageBuffer.contiguous();
        gstar_bwd_args a = {};
        a.P = P; a.D = degree; a.M = M; a.R = R;
        a.background = fptr(bg); a.width = W; a.height = H;
        a.means3D = fptr(m3); a.shs = fptr(shc); a.colors_precomp = fptr(col); a.scales = fptr(sc); a.scale_modifier = scale_modifier;
        a.rotations = fptr(rot); a.cov3D_precomp = fptr(cov); a.viewmatrix = fptr(vm); a.projmatrix = fptr(pm); a.campos = fptr(cp);
        a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
        a.radii = rad.data_ptr<int>();
        a.geom_buffer = reinterpret_cast<char*>(gb.data_ptr());
        a.binning_buffer = bb.numel() ? reinterpret_cast<char*>(bb.data_ptr()) : nullptr;
        a.image_buffer = reinterpret_cast<char*>(ib.data_ptr());
        a.dL_dpix = fptr(dpix);
        a.dL_dmean2D = dL_dmeans2D.data_ptr<float>(); a.dL_dconic = nullptr;
        a.dL_dopacity = dL_dopacity.data_ptr<float>(); a.dL_dcolor = has_colors ? dL_dcolors.data_ptr<float>() : nullptr;
        a.dL_dmean3D = dL_dmeans3D.data_ptr<float>(); a.dL_dcov3D = has_cov ? dL_dcov3D.data_ptr<float>() : nullptr;
        a.dL_dsh = M ? dL_dsh.data_ptr<float>() : nullptr;
        a.dL_dscale = dL_dscales.data_ptr<float>(); a.dL_drot = dL_drotations.data_ptr<float>();
        a.blend_grad_scratch = scratch.data_ptr<float>();
        a.debug = debug ? 1 : 0;
        a.accumulate_param_grads = fused ? (fused->atomic ? 2 : 1) : 0;
        a.blend_only = 0;
        torch::Tensor dpix2, bg2, col2;
        if (second) {
            dpix2 = prep(second->dL_dout_color2, dev); bg2 = prep(second->background2, dev); col2 = pad4(prep(second->colors2, dev));
            a.dL_dpix2 = fptr(dpix2); a.background2 = fptr(bg2); a.colors2 = fptr(col2); a.channels2 = (int)dpix2.size(0);
            if (a.channels2 == 4) {
                TORCH_CHECK(second->scratch2.is_cuda() && second->scratch2.scalar_type() == torch::kFloat32 && second->scratch2.is_contiguous() &&
                                second->scratch2.numel() == P,
                            "gaustar_b200: a four-channel second pass needs a zeroed float32 CUDA scratch of P values for the fourth colour moment");
                a.blend_grad_scratch2 = second->scratch2.data_ptr<float>();
            }
        }
        check(gstar_raster_backward(&a, stream));
    }
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
}

static void check_scratch(const torch::Tensor& scratch, const torch::Tensor& like, int64_t P)
{
    TORCH_CHECK(scratch.is_cuda() && scratch.device() == like.device() && scratch.scalar_type() == torch::kFloat32 && scratch.is_contiguous() &&
                    scratch.dim() == 2 && scratch.size(0) == P && scratch.size(1) == GSTAR_GRAD_SCRATCH_FLOATS,
                "gaustar_b200: the moment scratch must be a contiguous float32 CUDA tensor of shape (P, 12)");
}

// Several feature passes over one geometry (gstar_bwd_args.blend_only): ADD the blend-stage moments of one pass into
// `scratch` (P x 12); nothing else is computed.  Columns 6..8 are then that pass's dL_dcolors.
void BlendBackwardCUDA(const torch::Tensor& background, const torch::Tensor& dL_dout_color, const torch::Tensor& radii,
                       const torch::Tensor& geomBuffer, const int R, const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
                       torch::Tensor scratch, const bool debug)
{
    TORCH_CHECK(dL_dout_color.is_cuda() && dL_dout_color.dim() == 3, "gaustar_b200: dL_dout_color must be a CUDA tensor of shape (3, H, W)");
    const torch::Device dev = dL_dout_color.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t P = scratch.size(0);
    check_scratch(scratch, dL_dout_color, P);
    if (P == 0) return;
    const torch::Tensor bg = prep(background, dev), dpix = prep(dL_dout_color, dev);
    const torch::Tensor rad = radii.contiguous(), gb = geomBuffer.contiguous(), bb = binningBuffer.contiguous(), ib = imageBuffer.contiguous();
    gstar_bwd_args a = {};
    a.P = (int)P; a.R = R;
    a.background = fptr(bg); a.width = (int)dL_dout_color.size(2); a.height = (int)dL_dout_color.size(1);
    a.radii = rad.data_ptr<int>();
    a.geom_buffer = reinterpret_cast<char*>(gb.data_ptr());
    a.binning_buffer = bb.numel() ? reinterpret_cast<char*>(bb.data_ptr()) : nullptr;
    a.image_buffer = reinterpret_cast<char*>(ib.data_ptr());
    a.dL_dpix = fptr(dpix);
    a.blend_grad_scratch = scratch.data_ptr<float>();
    a.debug = debug ? 1 : 0;
    a.blend_only = 1;
    check(gstar_raster_backward(&a, c10::cuda::getCurrentCUDAStream(dev.index()).stream()));
}

// The full backward of one pass on a scratch that blend-only calls of other passes pre-loaded (same 21 arguments as
// rasterize_gaussians_backward + the scratch).
BwdTuple RasterizeGaussiansBackwardPreloadedCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                                                 const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                                                 const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                                                 const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                                                 const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree,
                                                 const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                                                 const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const bool debug,
                                                 torch::Tensor scratch)
{
    check_scratch(scratch, means3D, means3D.size(0));
    return backward_impl(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                         tan_fovy, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug, nullptr, &scratch);
}

// Backward of rasterize_gaussians_dual: the 21 arguments of rasterize_gaussians_backward + the moment scratch (P x 12, zero or
// pre-loaded by blend-only calls of further passes) + the second image's upstream gradient (c, H, W), background (c) and colours
// (P, c), c = 1..4, and a zeroed scratch of P floats (used iff c == 4; otherwise an empty tensor).  Afterwards columns 9..11 of the
// scratch are the dL_dcolors of the second pass's first three channels and scratch2 that of the fourth.
BwdTuple RasterizeGaussiansBackwardDualCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                                            const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                                            const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                                            const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                                            const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree,
                                            const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                                            const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const bool debug,
                                            torch::Tensor scratch, const torch::Tensor& dL_dout_color2, const torch::Tensor& background2,
                                            const torch::Tensor& colors2, torch::Tensor scratch2)
{
    check_scratch(scratch, means3D, means3D.size(0));
    TORCH_CHECK(dL_dout_color2.is_cuda() && dL_dout_color2.dim() == 3 && dL_dout_color2.size(1) == dL_dout_color.size(1) &&
                    dL_dout_color2.size(2) == dL_dout_color.size(2) && dL_dout_color2.size(0) == colors2.size(1),
                "gaustar_b200: the second upstream gradient must be (c, H, W) with c the second pass's channels");
    const SecondPass sp{dL_dout_color2, background2, colors2, scratch2};
    return backward_impl(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                         tan_fovy, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer, debug, nullptr, &scratch, &sp);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix)
{
    TORCH_CHECK(means3D.is_cuda(), "gaustar_b200: means3D must be a CUDA tensor (there is no CPU path)");
    const torch::Device dev = means3D.device();
    c10::cuda::CUDAGuard guard(dev);
    const int P = means3D.size(0);
    torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
    if (P != 0) {
        const torch::Tensor m3 = prep(means3D, dev), vm = prep(viewmatrix, dev), pm = prep(projmatrix, dev);
        check(gstar_mark_visible(P, fptr(m3), fptr(vm), fptr(pm), reinterpret_cast<unsigned char*>(present.data_ptr<bool>()),
                                 c10::cuda::getCurrentCUDAStream(dev.index()).stream()));
    }
    return present;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
    m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
    m.def("rasterize_gaussians_backward_fused", &RasterizeGaussiansBackwardFusedCUDA);
    m.def("rasterize_gaussians_reblend", &ReblendGaussiansCUDA);
    m.def("rasterize_gaussians_blend_backward", &BlendBackwardCUDA);
    m.def("rasterize_gaussians_backward_preloaded", &RasterizeGaussiansBackwardPreloadedCUDA);
    m.def("rasterize_gaussians_dual", &RasterizeGaussiansDualCUDA);
    m.def("rasterize_gaussians_backward_dual", &RasterizeGaussiansBackwardDualCUDA);
    m.def("mark_visible", &markVisible);
    m.def("set_forward_only", &set_forward_only);
}
