// compute_shader.cpp -- see compute_shader.h.  Error convention of gdcs:
// UtilityFunctions::printerr + early return (gdcs.cpp:22-67), compute() is a
// silent no-op when check_ready() is false (gdcs.cpp:239-240).
#include "compute_shader.h"

#include <cstdio>

namespace gdpt {

void ComputeShader::report(const char *what) const
{
    std::fprintf(stderr, "gdpt::ComputeShader: %s: %s\n", what, gdpt_last_error(rd_));
}

ComputeShader::ComputeShader(const std::string &shader_path, gdpt_device *rd, const std::vector<std::string> &args, int cuda_ordinal)
{
    rd_ = rd;
    if (!rd_) {
        if (gdpt_device_create(cuda_ordinal, &rd_) != GDPT_OK) {
            std::fprintf(stderr, "Failed to create rendering device: %s\n", gdpt_last_error(nullptr));
            rd_ = nullptr;
            return;
        }
        owns_rd_ = true;
    }
    std::vector<const char *> argv;
    for (const std::string &a : args) argv.push_back(a.c_str());
    if (gdpt_shader_create(rd_, shader_path.c_str(), argv.data(), (int)argv.size(), &shader_) != GDPT_OK) {
        report("Failed to load shader file");
        shader_ = nullptr;
    }
}

ComputeShader::~ComputeShader()
{
    if (shader_) gdpt_shader_destroy(shader_);
    if (owns_rd_ && rd_) gdpt_device_destroy(rd_);
}

gdpt_rid ComputeShader::create_storage_buffer_uniform(const void *data, uint64_t size, int binding, int set)
{
    if (!shader_) return 0;
    const gdpt_rid rid = gdpt_shader_create_storage_buffer_uniform(shader_, data, size, binding, set);
    if (!rid) report("create_storage_buffer_uniform");
    return rid;
}

void ComputeShader::update_storage_buffer_uniform(gdpt_rid rid, const void *data, uint64_t size)
{
    if (!shader_) return;
    if (gdpt_shader_update_storage_buffer_uniform(shader_, rid, data, size) != GDPT_OK) report("update_storage_buffer_uniform");
}

std::vector<uint8_t> ComputeShader::get_storage_buffer_uniform(gdpt_rid rid, uint64_t size) const
{
    std::vector<uint8_t> out(size);
    if (!shader_ || gdpt_shader_get_storage_buffer_uniform(shader_, rid, out.data(), size) != GDPT_OK) {
        if (shader_) report("get_storage_buffer_uniform");
        out.clear();
    }
    return out;
}

gdpt_rid ComputeShader::create_image_uniform(const void *pixels, int width, int height, gdpt_data_format format, int binding, int set)
{
    if (!shader_) return 0;
    const gdpt_rid rid = gdpt_shader_create_image_uniform(shader_, pixels, width, height, format, binding, set);
    if (!rid) report("create_image_uniform");
    return rid;
}

std::vector<uint8_t> ComputeShader::get_image_uniform_buffer(gdpt_rid rid, uint64_t layer_bytes, int layer) const
{
    std::vector<uint8_t> out(layer_bytes);
    if (get_image_uniform_buffer_into(rid, out.data(), layer_bytes, layer) != GDPT_OK) out.clear();
    return out;
}

int ComputeShader::get_image_uniform_buffer_into(gdpt_rid rid, void *out, uint64_t capacity, int layer) const
{
    if (!shader_) return GDPT_ERR_NOT_READY;
    const int rc = gdpt_shader_get_image_uniform_buffer(shader_, rid, layer, out, capacity);
    if (rc != GDPT_OK) report("get_image_uniform_buffer");
    return rc;
}

gdpt_rid ComputeShader::create_layered_image_uniform(const std::vector<const void *> &layers, int width, int height,
                                                     gdpt_data_format format, int binding, int set)
{
    if (!shader_) return 0;
    const gdpt_rid rid = gdpt_shader_create_layered_image_uniform(shader_, layers.data(), (int)layers.size(), width, height, format, binding, set);
    if (!rid) report("create_layered_image_uniform");
    return rid;
}

void ComputeShader::add_existing_buffer(gdpt_rid rid, gdpt_uniform_type uniform_type, int binding, int set)
{
    if (!shader_) return;
    if (gdpt_shader_add_existing_buffer(shader_, rid, uniform_type, binding, set) != GDPT_OK) report("add_existing_buffer");
}

void ComputeShader::finish_create_uniforms()
{
    if (!shader_) return;
    if (gdpt_shader_finish_create_uniforms(shader_) != GDPT_OK) report("finish_create_uniforms");
}

bool ComputeShader::check_ready() const { return shader_ && gdpt_shader_check_ready(shader_) == 1; }

void ComputeShader::compute(int gx, int gy, int gz)
{
    if (!check_ready()) return;
    if (gdpt_shader_compute(shader_, gx, gy, gz) != GDPT_OK) report("compute");
}

} // namespace gdpt
