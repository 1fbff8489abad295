// gdcs_cuda.cpp -- the reference's gdcs `ComputeShader` (src/gdcs/include/gdcs.h:21-72, UNCHANGED) implemented over
// libgdpt_cuda.so instead of Godot's RenderingDevice.  Drop-in replacement for src/gdcs/src/gdcs.cpp: with this file
// in its place (and -lgdpt_cuda on the link line of SConstruct) PathTracingCamera (src/path_tracing/
// path_tracing_camera.cpp:111-232), ProgressiveRendering and TemporalReprojection (src/path_tracing/post_processing/)
// compile and run unmodified; every method below replaces the one of the same name in gdcs.cpp (lines cited).
//
// gdcs.h may not change, so the CUDA side lives in the members it already has:
//   _rd        the RenderingDevice the caller handed in.  It is only an identity here: every ComputeShader created
//              with the same `rd` shares one gdpt_device (path_tracing_camera.cpp:114 creates the device once and gives
//              it to the main shader and to both post processes, progressive_rendering.cpp:25).
//   _shader    the gdpt_shader handle, _pipeline the gdpt_device handle (RIDs are 64-bit ids; both are pointers).
//   _buffers   RIDs this shader created (freed with it, as upstream).
// Error convention as upstream: print and carry on; a shader that failed to build stays un-ready and compute() is a
// no-op (gdcs.cpp:22-67, 239-240).
//
// Compile check (no Godot binary needed): tests/test_adapter_compile.py generates the godot-cpp headers offline and
// runs g++ -fsyntax-only on this file and on the reference's own, unmodified callers.
#include "gdcs.h"

#include <godot_cpp/classes/image.hpp>
#include <godot_cpp/classes/rd_shader_source.hpp>
#include <godot_cpp/variant/utility_functions.hpp>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "gdpt.h"

namespace {

struct SharedDevice { gdpt_device *dev = nullptr; int users = 0; };
std::mutex g_lock;
std::unordered_map<RenderingDevice *, SharedDevice> g_devices; // one CUDA device object per RenderingDevice identity

gdpt_device *acquire_device(RenderingDevice *rd)
{
    std::lock_guard<std::mutex> hold(g_lock);
    SharedDevice &s = g_devices[rd];
    if (s.dev == nullptr && gdpt_device_create(0, &s.dev) != GDPT_OK) {
        UtilityFunctions::printerr("Failed to create CUDA device: ", gdpt_last_error(nullptr));
        g_devices.erase(rd);
        return nullptr;
    }
    s.users++;
    return s.dev;
}

// true when the caller was the last user (then the RenderingDevice itself may go too, gdcs.cpp:86)
bool release_device(RenderingDevice *rd)
{
    std::lock_guard<std::mutex> hold(g_lock);
    auto it = g_devices.find(rd);
    if (it == g_devices.end()) return true;
    if (--it->second.users > 0) return false;
    gdpt_device_destroy(it->second.dev);
    g_devices.erase(it);
    return true;
}

RID to_rid(uint64_t v) { return UtilityFunctions::rid_from_int64((int64_t)v); }
uint64_t from_rid(const RID &r) { return (uint64_t)r.get_id(); }
gdpt_shader *shader_of(const RID &r) { return reinterpret_cast<gdpt_shader *>(from_rid(r)); }
gdpt_device *device_of(const RID &r) { return reinterpret_cast<gdpt_device *>(from_rid(r)); }

bool to_gdpt_format(RenderingDevice::DataFormat f, gdpt_data_format *out)
{
    switch (f) {
    case RenderingDevice::DATA_FORMAT_R8G8B8A8_UNORM: *out = GDPT_FORMAT_R8G8B8A8_UNORM; return true;      // path_tracing_camera.cpp:148,182
    case RenderingDevice::DATA_FORMAT_R32_SFLOAT: *out = GDPT_FORMAT_R32_SFLOAT; return true;              // path_tracing_camera.cpp:160
    case RenderingDevice::DATA_FORMAT_R32G32B32A32_SFLOAT: *out = GDPT_FORMAT_R32G32B32A32_SFLOAT; return true; // progressive_rendering.cpp:33
    default: return false;
    }
}

} // namespace

// gdcs.cpp:15-74
ComputeShader::ComputeShader(const String &shader_path, RenderingDevice *rd, const std::vector<String> args)
{
    _rd = rd != nullptr ? rd : RenderingServer::get_singleton()->create_local_rendering_device();
    gdpt_device *dev = acquire_device(_rd);
    if (dev == nullptr) return;
    _pipeline = to_rid(reinterpret_cast<uint64_t>(dev));
    // the "#define ..." strings gdcs injects into the GLSL source (gdcs.cpp:277-322) select the same options here
    std::vector<CharString> keep;
    std::vector<const char *> argv;
    keep.reserve(args.size());
    for (const String &a : args) {
        keep.push_back(a.utf8());
        argv.push_back(keep.back().get_data());
    }
    gdpt_shader *sh = nullptr;
    if (gdpt_shader_create(dev, shader_path.utf8().get_data(), argv.empty() ? nullptr : argv.data(), (int)argv.size(), &sh) != GDPT_OK) {
        UtilityFunctions::printerr("Failed to create shader: ", gdpt_last_error(dev));
        return;
    }
    _shader = to_rid(reinterpret_cast<uint64_t>(sh));
#ifdef GDCS_VERBOSE
    UtilityFunctions::print("loaded shader successfully!");
#endif
    _initialized = true;
}

// gdcs.cpp:76-87.  Upstream deletes `_rd` in every ComputeShader that shares it; here the last sharer does.
ComputeShader::~ComputeShader()
{
    if (_shader.is_valid()) gdpt_shader_destroy(shader_of(_shader)); // frees the RIDs this shader created
    _buffers.clear();
    if (release_device(_rd)) memdelete(_rd);
}

//------------------------------------------------ STORAGE BUFFER ------------------------------------------------

// gdcs.cpp:91-106
RID ComputeShader::create_storage_buffer_uniform(const PackedByteArray &data, const int binding, const int set)
{
    if (!_initialized) return RID();
    const gdpt_rid rid = gdpt_shader_create_storage_buffer_uniform(shader_of(_shader), data.ptr(), (uint64_t)data.size(), binding, set);
    if (rid == 0) UtilityFunctions::printerr("create_storage_buffer_uniform: ", gdpt_last_error(device_of(_pipeline)));
    _buffers.push_back(to_rid(rid));
    _uniforms_ready = false;
    return to_rid(rid);
}

// gdcs.cpp:108-111
void ComputeShader::update_storage_buffer_uniform(const RID rid, const PackedByteArray &data)
{
    if (!_initialized) return;
    if (gdpt_shader_update_storage_buffer_uniform(shader_of(_shader), from_rid(rid), data.ptr(), (uint64_t)data.size()) != GDPT_OK)
        UtilityFunctions::printerr("update_storage_buffer_uniform: ", gdpt_last_error(device_of(_pipeline)));
}

// gdcs.cpp:113-116
PackedByteArray ComputeShader::get_storage_buffer_uniform(RID rid) const
{
    PackedByteArray out;
    if (!_initialized) return out;
    uint64_t ptr = 0, size = 0;
    if (gdpt_rid_device_pointer(device_of(_pipeline), from_rid(rid), &ptr, &size) != GDPT_OK) return out;
    out.resize((int64_t)size);
    if (gdpt_shader_get_storage_buffer_uniform(shader_of(_shader), from_rid(rid), out.ptrw(), size) != GDPT_OK)
        UtilityFunctions::printerr("get_storage_buffer_uniform: ", gdpt_last_error(device_of(_pipeline)));
    return out;
}

//------------------------------------------------ TEXTURE 2D ------------------------------------------------

// gdcs.cpp:118-133: unchanged -- the format object is how the callers say width, height and data format
Ref<RDTextureFormat> ComputeShader::create_texture_format(const int width, const int height, const RenderingDevice::DataFormat format)
{
    Ref<RDTextureFormat> result;
    result.instantiate();
    result->set_width(width);
    result->set_height(height);
    result->set_format(format);
    result->set_usage_bits(RenderingDevice::TEXTURE_USAGE_STORAGE_BIT | RenderingDevice::TEXTURE_USAGE_CAN_UPDATE_BIT |
                           RenderingDevice::TEXTURE_USAGE_CAN_COPY_FROM_BIT);
    return result;
}

// gdcs.cpp:135-167
RID ComputeShader::create_image_uniform(const Ref<Image> &image, const Ref<RDTextureFormat> &format, const Ref<RDTextureView> &view,
                                        const int binding, const int set)
{
    if (!_initialized || format.is_null()) return RID();
    gdpt_data_format f;
    if (!to_gdpt_format(format->get_format(), &f)) {
        UtilityFunctions::printerr("create_image_uniform: data format not supported by the CUDA backend");
        return RID();
    }
    // upstream uploads image->get_data() when an image is given (gdcs.cpp:150-153), else an uninitialised texture
    const PackedByteArray pixels = image.is_valid() ? image->get_data() : PackedByteArray();
    const gdpt_rid rid = gdpt_shader_create_image_uniform(shader_of(_shader), pixels.size() ? pixels.ptr() : nullptr,
                                                         (int)format->get_width(), (int)format->get_height(), f, binding, set);
    if (rid == 0) UtilityFunctions::printerr("create_image_uniform: ", gdpt_last_error(device_of(_pipeline)));
    _buffers.push_back(to_rid(rid));
    _uniforms_ready = false;
    return to_rid(rid);
}

// gdcs.cpp:169-172
PackedByteArray ComputeShader::get_image_uniform_buffer(RID rid, const int layer) const
{
    PackedByteArray out;
    if (!_initialized) return out;
    uint64_t ptr = 0, size = 0;
    if (gdpt_rid_device_pointer(device_of(_pipeline), from_rid(rid), &ptr, &size) != GDPT_OK) return out;
    out.resize((int64_t)size); // one layer of a 2D image; the layered texture is never read back upstream
    if (gdpt_shader_get_image_uniform_buffer(shader_of(_shader), from_rid(rid), layer, out.ptrw(), size) != GDPT_OK)
        UtilityFunctions::printerr("get_image_uniform_buffer: ", gdpt_last_error(device_of(_pipeline)));
    return out;
}

// gdcs.cpp:174-207
RID ComputeShader::create_layered_image_uniform(const std::vector<Ref<Image>> &image, const Ref<RDTextureFormat> &format,
                                                const Ref<RDTextureView> &view, const int binding, const int set)
{
    if (!_initialized || format.is_null() || image.empty()) return RID();
    gdpt_data_format f;
    if (!to_gdpt_format(format->get_format(), &f)) {
        UtilityFunctions::printerr("create_layered_image_uniform: data format not supported by the CUDA backend");
        return RID();
    }
    std::vector<PackedByteArray> keep;
    std::vector<const void *> layers;
    keep.reserve(image.size());
    for (const Ref<Image> &im : image) {
        keep.push_back(im->get_data());
        layers.push_back(keep.back().ptr());
    }
    const gdpt_rid rid = gdpt_shader_create_layered_image_uniform(shader_of(_shader), layers.data(), (int)layers.size(),
                                                                 (int)format->get_width(), (int)format->get_height(), f, binding, set);
    if (rid == 0) UtilityFunctions::printerr("create_layered_image_uniform: ", gdpt_last_error(device_of(_pipeline)));
    _buffers.push_back(to_rid(rid));
    _uniforms_ready = false;
    return to_rid(rid);
}

//------------------------------------------------ GENERAL ------------------------------------------------

// gdcs.cpp:209-223
void ComputeShader::add_existing_buffer(const RID rid, const RenderingDevice::UniformType uniform_type, const int binding, const int set)
{
    if (!_initialized) return;
    const gdpt_uniform_type t = uniform_type == RenderingDevice::UNIFORM_TYPE_IMAGE ? GDPT_UNIFORM_TYPE_IMAGE : GDPT_UNIFORM_TYPE_STORAGE_BUFFER;
    if (gdpt_shader_add_existing_buffer(shader_of(_shader), from_rid(rid), t, binding, set) != GDPT_OK)
        UtilityFunctions::printerr("add_existing_buffer: ", gdpt_last_error(device_of(_pipeline)));
    _uniforms_ready = false;
}

// gdcs.cpp:225-235
void ComputeShader::finish_create_uniforms()
{
    if (!_initialized) return;
    if (gdpt_shader_finish_create_uniforms(shader_of(_shader)) != GDPT_OK) {
        UtilityFunctions::printerr("finish_create_uniforms: ", gdpt_last_error(device_of(_pipeline)));
        return;
    }
    _uniforms_ready = true;
}

// gdcs.cpp:258-273
bool ComputeShader::check_ready() const
{
    if (!_rd || !_initialized || !_uniforms_ready) return false;
    return gdpt_shader_check_ready(shader_of(_shader)) == 1;
}

// gdcs.cpp:277-354: the GLSL source is not compiled here; kept because gdcs.h declares it
Ref<RDShaderSource> ComputeShader::LoadShaderFile(const String &shader_path, const std::vector<String> &args) { return Ref<RDShaderSource>(); }
String ComputeShader::LoadShaderString(const String &shader_path) { return String(); }

// gdcs.cpp:237-251: dispatch + submit() + sync() = one blocking call
void ComputeShader::compute(const Vector3i groups)
{
    if (!check_ready()) return;
    if (gdpt_shader_compute(shader_of(_shader), groups.x, groups.y, groups.z) != GDPT_OK)
        UtilityFunctions::printerr("compute: ", gdpt_last_error(device_of(_pipeline)));
}

// gdcs.cpp:253-256
RenderingDevice *ComputeShader::get_rendering_device() const { return _rd; }
