// compute_shader.h -- host twin of gdcs `ComputeShader`
// (src/gdcs/include/gdcs.h:21-72): same method names, argument meaning and
// error behaviour (print to stderr + no-op, never throw), but every method
// forwards to the C-ABI of libgdpt_cuda.so instead of Godot's RenderingDevice.
// PackedByteArray becomes (pointer, size); Ref<Image> + RDTextureFormat become
// (pixels, width, height, format); RID becomes gdpt_rid.
#ifndef GDPT_COMPUTE_SHADER_H
#define GDPT_COMPUTE_SHADER_H

#include "gdpt.h"

#include <cstdint>
#include <string>
#include <vector>

namespace gdpt {

class ComputeShader {
public:
    // rd == nullptr: create (and own) a local device, like gdcs.cpp:17-20.
    ComputeShader(const std::string &shader_path, gdpt_device *rd = nullptr, const std::vector<std::string> &args = {},
                  int cuda_ordinal = 0);
    ~ComputeShader();
    ComputeShader(const ComputeShader &) = delete;
    ComputeShader &operator=(const ComputeShader &) = delete;

    // storage buffers
    gdpt_rid create_storage_buffer_uniform(const void *data, uint64_t size, int binding, int set = 0);
    void update_storage_buffer_uniform(gdpt_rid rid, const void *data, uint64_t size);
    std::vector<uint8_t> get_storage_buffer_uniform(gdpt_rid rid, uint64_t size) const;

    // 2d textures
    gdpt_rid create_image_uniform(const void *pixels, int width, int height, gdpt_data_format format, int binding, int set = 0);
    std::vector<uint8_t> get_image_uniform_buffer(gdpt_rid rid, uint64_t layer_bytes, int layer = 0) const;
    int get_image_uniform_buffer_into(gdpt_rid rid, void *out, uint64_t capacity, int layer = 0) const;

    // 2d layered textures
    gdpt_rid create_layered_image_uniform(const std::vector<const void *> &layers, int width, int height,
                                          gdpt_data_format format, int binding, int set = 0);

    // general
    void add_existing_buffer(gdpt_rid rid, gdpt_uniform_type uniform_type, int binding, int set = 0);
    void finish_create_uniforms();
    bool check_ready() const;
    void compute(int groups_x, int groups_y, int groups_z);

    gdpt_device *get_rendering_device() const { return rd_; }
    gdpt_shader *handle() const { return shader_; }

private:
    void report(const char *what) const;
    gdpt_device *rd_ = nullptr;
    bool owns_rd_ = false;
    gdpt_shader *shader_ = nullptr;
};

} // namespace gdpt
#endif
