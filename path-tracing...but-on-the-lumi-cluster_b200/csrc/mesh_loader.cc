// mesh_loader.cc — host-only: OBJ/MTL files -> the mesh buffers the render path consumes.
//
// Restates load_mesh / load_mtl (mesh.cc:52-265) behind ptgpu_meshes_*: same commands (v, vn, vt, f,
// usemtl, mtllib; newmtl, Kd, Ke, d, Pr, Pm, Tf), same vertex de-duplication (one vertex per distinct
// (position, texcoord, normal, material) index group, numbered by first appearance, mesh.cc:215-262),
// same attribute packing (albedo.w = alpha; material = roughness, metallicness, max transmission, max
// scaled emission, mesh.cc:230-250). Differences in mechanism, not in result: the file is tokenised in
// place, numbers are parsed with the locale-independent std::from_chars (the reference relies on
// main.cc:63 setting the C locale for strtof), and the index groups are de-duplicated with a hash map
// instead of a std::map. A missing file is an error code + message here; the reference exits
// (mesh.cc:25-29), which the C++ caller can still do.
#include "../../include/ptgpu.h"

#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Material
{
    std::string name;
    float albedo[3] = {1, 1, 1};
    float alpha = 0;
    float emission[3] = {0, 0, 0};
    float roughness = 1;
    float metallicness = 0;
    float transmission[3] = {0, 0, 0};
};

bool read_file(const std::string& path, std::string& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if(!f) return false;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(sz < 0 ? 0 : (size_t)sz);
    bool ok = sz >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

// strtof(str, &str) of the reference: skips white space, parses a number if there is one, else 0 and
// the cursor stays
float take_float(const char*& s, const char* end)
{
    const char* p = s;
    while(p < end && isspace((unsigned char)*p)) ++p;
    const char* q = p;
    if(q < end && *q == '+') ++q;
    float v = 0.0f;
    auto r = std::from_chars(q, end, v);
    if(r.ec == std::errc::invalid_argument) return 0.0f;
    if(r.ec == std::errc::result_out_of_range)
    {   // strtof saturates (HUGE_VALF / 0) instead of failing
        char* e = nullptr;
        v = strtof(std::string(p, r.ptr).c_str(), &e);
    }
    s = r.ptr;
    return v;
}

// strtol(str, &str, 0) of the reference
long take_long(const char*& s, const char* end)
{
    const char* p = s;
    while(p < end && isspace((unsigned char)*p)) ++p;
    bool neg = false;
    const char* q = p;
    if(q < end && (*q == '+' || *q == '-')) { neg = *q == '-'; ++q; }
    int base = 10;
    if(q + 1 < end && q[0] == '0' && (q[1] == 'x' || q[1] == 'X')) { base = 16; q += 2; }
    else if(q < end && q[0] == '0') base = 8;
    long v = 0;
    auto r = std::from_chars(q, end, v, base);
    if(r.ec == std::errc::invalid_argument)
    {
        if(base == 16) { s = q - 1; return 0; } // "0x" without digits: strtol consumes the "0"
        return 0;
    }
    s = r.ptr;
    return neg ? -v : v;
}

std::string take_word(const char*& s, const char* end)
{   // read_string, mesh.cc:43-50
    while(s < end && isspace((unsigned char)*s)) ++s;
    const char* b = s;
    while(s < end && !isspace((unsigned char)*s)) ++s;
    return std::string(b, s);
}

// The reference compares with strncmp(command, literal, command_len) == 0 (mesh.cc:70-101, 163-218): a
// command matches when it EQUALS the literal or is a proper prefix of it (the empty command after the
// last line matches the first literal tested). Kept as it is.
bool command_is(const char* cmd, size_t len, const char* literal)
{
    return strncmp(cmd, literal, len) == 0;
}

void skip_line(const char*& s, const char* end)
{
    while(s < end && *s != '\n') ++s;
}

void load_mtl(std::vector<Material>& materials, const std::string& path, std::string& err)
{
    std::string data;
    if(!read_file(path, data)) { err = "Unable to open " + path; return; }
    data.push_back('\0'); // commands are compared with strncmp: keep a terminator behind the text
    const char* s = data.data();
    const char* end = s + data.size() - 1;
    long cur = -1;
    while(s < end)
    {
        while(s < end && isspace((unsigned char)*s)) ++s;
        const char* cmd = s;
        while(s < end && !isspace((unsigned char)*s)) ++s;
        const size_t len = (size_t)(s - cmd);
        if(command_is(cmd, len, "newmtl"))
        {
            Material m;
            m.name = take_word(s, end);
            materials.push_back(m);
            cur = (long)materials.size() - 1;
        }
        else if(cur < 0) {}
        else if(command_is(cmd, len, "Kd")) { for(int k = 0; k < 3; ++k) materials[cur].albedo[k] = take_float(s, end); }
        else if(command_is(cmd, len, "Ke")) { for(int k = 0; k < 3; ++k) materials[cur].emission[k] = take_float(s, end); }
        else if(command_is(cmd, len, "d")) materials[cur].alpha = take_float(s, end);
        else if(command_is(cmd, len, "Pr")) materials[cur].roughness = take_float(s, end);
        else if(command_is(cmd, len, "Pm")) materials[cur].metallicness = take_float(s, end);
        else if(command_is(cmd, len, "Tf")) { for(int k = 0; k < 3; ++k) materials[cur].transmission[k] = take_float(s, end); }
        skip_line(s, end);
    }
}

struct IndexGroup
{
    int32_t pos, tex, normal, material;
    bool operator==(const IndexGroup& o) const { return pos == o.pos && tex == o.tex && normal == o.normal && material == o.material; }
};
struct IndexGroupHash
{
    size_t operator()(const IndexGroup& g) const
    {
        uint64_t h = (uint64_t)(uint32_t)g.pos * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)(uint32_t)g.tex + 0x7F4A7C15u) * 0xC2B2AE3D27D4EB4Full;
        h ^= ((uint64_t)(uint32_t)g.normal << 21) * 0x165667B19E3779F9ull;
        h ^= (uint64_t)(uint32_t)g.material * 0x27D4EB2F165667C5ull;
        return (size_t)(h ^ (h >> 29));
    }
};

} // namespace

struct ptgpu_mesh_set
{
    std::vector<uint32_t> indices;
    std::vector<ptgpu_float3> pos, normal;
    std::vector<ptgpu_float4> albedo, material;
    std::string error;
};

extern "C" {

int ptgpu_meshes_create(ptgpu_mesh_set** out)
{
    if(!out) return 1;
    *out = new ptgpu_mesh_set();
    return 0;
}

void ptgpu_meshes_destroy(ptgpu_mesh_set* set) { delete set; }

const char* ptgpu_meshes_last_error(const ptgpu_mesh_set* set) { return set ? set->error.c_str() : "null mesh set"; }

int ptgpu_meshes_load_obj(ptgpu_mesh_set* set, const char* obj_path, ptgpu_mesh* out_mesh)
{
    if(!set || !obj_path || !out_mesh) return 1;
    set->error.clear();
    std::string data;
    if(!read_file(obj_path, data)) { set->error = std::string("Unable to open ") + obj_path; return 1; }
    data.push_back('\0');
    const char* slash = strrchr(obj_path, '/');
    const std::string prefix = slash ? std::string(obj_path, slash + 1) : std::string(); // mesh.cc:145 (a path without '/' is undefined there)

    std::vector<ptgpu_float3> positions, normals;
    std::vector<Material> materials(1);          // material 0: the defaults (mesh.cc:147-148)
    std::vector<IndexGroup> groups;
    int32_t active_material = 0;

    const char* s = data.data();
    const char* end = s + data.size() - 1;
    while(s < end)
    {
        while(s < end && isspace((unsigned char)*s)) ++s;
        const char* cmd = s;
        while(s < end && !isspace((unsigned char)*s)) ++s;
        const size_t len = (size_t)(s - cmd);
        if(command_is(cmd, len, "v"))
        {
            ptgpu_float3 p{};
            p.x = take_float(s, end); p.y = take_float(s, end); p.z = take_float(s, end);
            positions.push_back(p);
        }
        else if(command_is(cmd, len, "vn"))
        {
            ptgpu_float3 n{};
            n.x = take_float(s, end); n.y = take_float(s, end); n.z = take_float(s, end);
            const float l = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z); // normalize, math.hh:106-110
            n.x /= l; n.y /= l; n.z /= l;
            normals.push_back(n);
        }
        else if(command_is(cmd, len, "vt")) { take_float(s, end); take_float(s, end); } // parsed, never used (no textures)
        else if(command_is(cmd, len, "f"))
        {
            for(int i = 0; i < 3; ++i)
            {   // triangles only (mesh.hh:46-49)
                IndexGroup g;
                g.material = active_material;
                g.pos = (int32_t)(take_long(s, end) - 1);
                if(s < end && *s == '/') ++s;
                g.tex = (int32_t)(take_long(s, end) - 1);
                if(s < end && *s == '/') ++s;
                g.normal = (int32_t)(take_long(s, end) - 1);
                groups.push_back(g);
            }
        }
        else if(command_is(cmd, len, "usemtl"))
        {
            const std::string name = take_word(s, end);
            for(size_t i = 0; i < materials.size(); ++i)
                if(materials[i].name == name) { active_material = (int32_t)i; break; }
        }
        else if(command_is(cmd, len, "mtllib"))
        {
            std::string err;
            load_mtl(materials, prefix + take_word(s, end), err);
            if(!err.empty()) { set->error = err; return 1; }
        }
        skip_line(s, end);
    }

    ptgpu_mesh m;
    m.index_offset = (uint32_t)set->indices.size();
    m.base_vertex_offset = (uint32_t)set->pos.size();
    m.triangle_count = (uint32_t)(groups.size() / 3);
    m.vertex_count = 0;
    std::unordered_map<IndexGroup, uint32_t, IndexGroupHash> seen;
    seen.reserve(groups.size());
    set->indices.reserve(set->indices.size() + groups.size());
    for(const IndexGroup& g : groups)
    {
        auto it = seen.find(g);
        if(it == seen.end())
        {
            it = seen.emplace(g, (uint32_t)seen.size()).first;
            ptgpu_float3 p{}, n{};
            if(g.pos >= 0 && (size_t)g.pos < positions.size()) p = positions[g.pos];
            if(g.normal >= 0 && (size_t)g.normal < normals.size()) n = normals[g.normal];
            ptgpu_float4 a{}, mt{};
            if(g.material >= 0 && (size_t)g.material < materials.size())
            {
                const Material& mat = materials[g.material];
                a.x = mat.albedo[0]; a.y = mat.albedo[1]; a.z = mat.albedo[2]; a.w = mat.alpha;
                mt.x = mat.roughness;
                mt.y = mat.metallicness;
                float se[3];
                for(int k = 0; k < 3; ++k)
                {   // emission / max(albedo, emission), clamped at 0; exactly 0 where the emission is 0 (mesh.cc:238-244)
                    se[k] = std::fmax(mat.emission[k] / std::fmax(mat.albedo[k], mat.emission[k]), 0.0f);
                    if(mat.emission[k] == 0) se[k] = 0;
                }
                mt.z = std::fmax(mat.transmission[0], std::fmax(mat.transmission[1], mat.transmission[2]));
                mt.w = std::fmax(se[0], std::fmax(se[1], se[2]));
            }
            set->pos.push_back(p); set->normal.push_back(n);
            set->albedo.push_back(a); set->material.push_back(mt);
            m.vertex_count++;
        }
        set->indices.push_back(it->second);
    }
    *out_mesh = m;
    return 0;
}

size_t ptgpu_meshes_index_count(const ptgpu_mesh_set* set) { return set ? set->indices.size() : 0; }
size_t ptgpu_meshes_vertex_count(const ptgpu_mesh_set* set) { return set ? set->pos.size() : 0; }
const uint32_t* ptgpu_meshes_indices(const ptgpu_mesh_set* set) { return set ? set->indices.data() : nullptr; }
const ptgpu_float3* ptgpu_meshes_pos(const ptgpu_mesh_set* set) { return set ? set->pos.data() : nullptr; }
const ptgpu_float3* ptgpu_meshes_normal(const ptgpu_mesh_set* set) { return set ? set->normal.data() : nullptr; }
const ptgpu_float4* ptgpu_meshes_albedo(const ptgpu_mesh_set* set) { return set ? set->albedo.data() : nullptr; }
const ptgpu_float4* ptgpu_meshes_material(const ptgpu_mesh_set* set) { return set ? set->material.data() : nullptr; }

} // extern "C"
