// onnx_wire.cpp — protobuf wire-format walk over ModelProto/GraphProto/NodeProto/AttributeProto/
// TensorProto/ValueInfoProto.  Field numbers per onnx.proto3, cross-checked
// against the reference fixture models/test_model/1/model.onnx.
#include "onnx_wire.h"

#include <cstring>
#include <fstream>
#include <stdexcept>

namespace b200 {
namespace onnx {
namespace {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool Done() const { return p >= end; }
    uint64_t Varint() {
        uint64_t v = 0;
        int shift = 0;
        while (true) {
            if (p >= end) throw std::runtime_error("onnx: truncated varint");
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
            if (shift > 63) throw std::runtime_error("onnx: varint too long");
        }
    }
    Reader Sub() {
        uint64_t n = Varint();
        if ((uint64_t)(end - p) < n) throw std::runtime_error("onnx: truncated length-delimited field");
        Reader r{p, p + n};
        p += n;
        return r;
    }
    void Skip(int wt) {
        switch (wt) {
            case 0: Varint(); break;
            case 1: Need(8); p += 8; break;
            case 2: Sub(); break;
            case 5: Need(4); p += 4; break;
            default: throw std::runtime_error("onnx: unsupported wire type");
        }
    }
    void Need(size_t n) {
        if ((size_t)(end - p) < n) throw std::runtime_error("onnx: truncated fixed field");
    }
    float Fixed32f() {
        Need(4);
        float f;
        memcpy(&f, p, 4);
        p += 4;
        return f;
    }
    double Fixed64d() {
        Need(8);
        double d;
        memcpy(&d, p, 8);
        p += 8;
        return d;
    }
    std::string Str() {
        Reader r = Sub();
        return std::string((const char*)r.p, (size_t)(r.end - r.p));
    }
};

// repeated int64, packed or not
void ReadInts(Reader& r, int wt, std::vector<int64_t>& out) {
    if (wt == 0) {
        out.push_back((int64_t)r.Varint());
    } else if (wt == 2) {
        Reader s = r.Sub();
        while (!s.Done()) out.push_back((int64_t)s.Varint());
    } else {
        r.Skip(wt);
    }
}
void ReadFloats(Reader& r, int wt, std::vector<float>& out) {
    if (wt == 5) {
        out.push_back(r.Fixed32f());
    } else if (wt == 2) {
        Reader s = r.Sub();
        while (!s.Done()) out.push_back(s.Fixed32f());
    } else {
        r.Skip(wt);
    }
}

float HalfToFloat(uint16_t h) {
    uint32_t sign = (h >> 15) & 1, exp = (h >> 10) & 0x1F, man = h & 0x3FF, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign << 31;
        } else {
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400));
            bits = (sign << 31) | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FF) << 13);
        }
    } else if (exp == 31) {
        bits = (sign << 31) | 0x7F800000u | (man << 13);
    } else {
        bits = (sign << 31) | ((exp + 112) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

TensorConst ParseTensor(Reader r) {
    TensorConst t;
    const uint8_t* raw = nullptr;
    size_t raw_n = 0;
    std::vector<float> fl;
    std::vector<double> dbl;
    std::vector<int64_t> i32, i64;
    while (!r.Done()) {
        uint64_t key = r.Varint();
        int fno = (int)(key >> 3), wt = (int)(key & 7);
        switch (fno) {
            case 1: ReadInts(r, wt, t.dims); break;
            case 2: t.dtype = (int)r.Varint(); break;
            case 4: ReadFloats(r, wt, fl); break;
            case 5: ReadInts(r, wt, i32); break;
            case 7: ReadInts(r, wt, i64); break;
            case 10:  // double_data: repeated fixed64, packed or not
                if (wt == 1) {
                    dbl.push_back(r.Fixed64d());
                } else if (wt == 2) {
                    Reader s = r.Sub();
                    while (!s.Done()) dbl.push_back(s.Fixed64d());
                } else {
                    r.Skip(wt);
                }
                break;
            case 8: t.name = r.Str(); break;
            case 9: {
                Reader s = r.Sub();
                raw = s.p;
                raw_n = (size_t)(s.end - s.p);
                break;
            }
            case 13: case 14:
                if (fno == 14 && wt == 0) {
                    if (r.Varint() == 1) throw std::runtime_error("onnx: external tensor data is not supported");
                } else {
                    r.Skip(wt);
                }
                break;
            default: r.Skip(wt);
        }
    }
    size_t n = 1;
    for (int64_t d : t.dims) {
        if (d < 0) throw std::runtime_error("onnx: tensor " + t.name + " has a negative dimension");
        if (d != 0 && n > (size_t)1 << 40) throw std::runtime_error("onnx: tensor " + t.name + " is implausibly large");
        n *= (size_t)d;
    }
    auto need = [&](size_t esz) {
        if (raw_n != n * esz) throw std::runtime_error("onnx: raw_data size mismatch for tensor " + t.name);
    };
    switch (t.dtype) {
        case kFloat:
            if (raw) { need(4); t.f32.resize(n); memcpy(t.f32.data(), raw, raw_n); }
            else t.f32 = fl;
            break;
        case kDouble:
            if (raw) {
                need(8); t.f32.resize(n);
                for (size_t i = 0; i < n; ++i) { double d; memcpy(&d, raw + 8 * i, 8); t.f32[i] = (float)d; }
            } else {
                t.f32.resize(dbl.size());
                for (size_t i = 0; i < dbl.size(); ++i) t.f32[i] = (float)dbl[i];
            }
            break;
        case kFloat16:
            t.f32.resize(n);
            if (raw) { need(2); for (size_t i = 0; i < n; ++i) { uint16_t h; memcpy(&h, raw + 2 * i, 2); t.f32[i] = HalfToFloat(h); } }
            else for (size_t i = 0; i < n && i < i32.size(); ++i) t.f32[i] = HalfToFloat((uint16_t)i32[i]);
            break;
        case kInt64:
            if (raw) { need(8); t.i64.resize(n); memcpy(t.i64.data(), raw, raw_n); }
            else t.i64 = i64;
            break;
        case kInt32:
            if (raw) { need(4); t.i64.resize(n); for (size_t i = 0; i < n; ++i) { int32_t v; memcpy(&v, raw + 4 * i, 4); t.i64[i] = v; } }
            else t.i64 = i32;
            break;
        case kUint8: case kInt8: case kBool:
            if (raw) { need(1); t.i64.resize(n); for (size_t i = 0; i < n; ++i) t.i64[i] = t.dtype == kInt8 ? (int8_t)raw[i] : raw[i]; }
            else t.i64 = i32;
            break;
        default:
            throw std::runtime_error("onnx: unsupported tensor data type " + std::to_string(t.dtype));
    }
    // every tensor must carry exactly the elements its dims announce (an empty payload for n > 0 used to slip through and
    // was indexed later)
    const bool is_float = t.dtype == kFloat || t.dtype == kDouble || t.dtype == kFloat16;
    if ((is_float ? t.f32.size() : t.i64.size()) != n) throw std::runtime_error("onnx: element count mismatch for tensor " + t.name);
    return t;
}

Attr ParseAttr(Reader r, std::string* name) {
    Attr a;
    bool has_f = false, has_i = false, has_s = false, has_t = false;
    while (!r.Done()) {
        uint64_t key = r.Varint();
        int fno = (int)(key >> 3), wt = (int)(key & 7);
        switch (fno) {
            case 1: *name = r.Str(); break;
            case 2: a.f = r.Fixed32f(); has_f = true; break;
            case 3: a.i = (int64_t)r.Varint(); has_i = true; break;
            case 4: a.s = r.Str(); has_s = true; break;
            case 5: a.t = ParseTensor(r.Sub()); has_t = true; break;
            case 7: ReadFloats(r, wt, a.floats); break;
            case 8: ReadInts(r, wt, a.ints); break;
            case 20: a.type = (int)r.Varint(); break;
            default: r.Skip(wt);
        }
    }
    if (a.type == 0) a.type = has_f ? 1 : has_i ? 2 : has_s ? 3 : has_t ? 4 : !a.floats.empty() ? 6 : 7;
    return a;
}

Node ParseNode(Reader r) {
    Node n;
    while (!r.Done()) {
        uint64_t key = r.Varint();
        int fno = (int)(key >> 3), wt = (int)(key & 7);
        switch (fno) {
            case 1: n.inputs.push_back(r.Str()); break;
            case 2: n.outputs.push_back(r.Str()); break;
            case 3: n.name = r.Str(); break;
            case 4: n.op_type = r.Str(); break;
            case 5: { std::string k; Attr a = ParseAttr(r.Sub(), &k); n.attrs[k] = std::move(a); break; }
            default: r.Skip(wt);
        }
    }
    return n;
}

ValueInfo ParseValueInfo(Reader r) {
    ValueInfo vi;
    while (!r.Done()) {
        uint64_t key = r.Varint();
        int fno = (int)(key >> 3), wt = (int)(key & 7);
        if (fno == 1) { vi.name = r.Str(); continue; }
        if (fno != 2) { r.Skip(wt); continue; }
        Reader tp = r.Sub();  // TypeProto
        while (!tp.Done()) {
            uint64_t k2 = tp.Varint();
            if ((k2 >> 3) != 1) { tp.Skip((int)(k2 & 7)); continue; }
            Reader tt = tp.Sub();  // TypeProto.Tensor
            while (!tt.Done()) {
                uint64_t k3 = tt.Varint();
                int f3 = (int)(k3 >> 3);
                if (f3 == 1) { vi.elem_type = (int)tt.Varint(); continue; }
                if (f3 != 2) { tt.Skip((int)(k3 & 7)); continue; }
                Reader sh = tt.Sub();  // TensorShapeProto
                while (!sh.Done()) {
                    uint64_t k4 = sh.Varint();
                    if ((k4 >> 3) != 1) { sh.Skip((int)(k4 & 7)); continue; }
                    Reader dim = sh.Sub();
                    int64_t v = -1;
                    while (!dim.Done()) {
                        uint64_t k5 = dim.Varint();
                        if ((k5 >> 3) == 1) v = (int64_t)dim.Varint();
                        else dim.Skip((int)(k5 & 7));
                    }
                    vi.dims.push_back(v);
                }
            }
        }
    }
    return vi;
}

Graph ParseGraph(Reader r) {
    Graph g;
    std::vector<ValueInfo> raw_inputs;
    while (!r.Done()) {
        uint64_t key = r.Varint();
        int fno = (int)(key >> 3), wt = (int)(key & 7);
        switch (fno) {
            case 1: g.nodes.push_back(ParseNode(r.Sub())); break;
            case 2: g.name = r.Str(); break;
            case 5: { TensorConst t = ParseTensor(r.Sub()); std::string nm = t.name; g.initializers[nm] = std::move(t); break; }
            case 11: raw_inputs.push_back(ParseValueInfo(r.Sub())); break;
            case 12: g.outputs.push_back(ParseValueInfo(r.Sub())); break;
            default: r.Skip(wt);
        }
    }
    for (auto& vi : raw_inputs)
        if (!g.initializers.count(vi.name)) g.inputs.push_back(vi);
    return g;
}

}  // namespace

int64_t Node::GetInt(const std::string& k, int64_t dflt) const {
    auto it = attrs.find(k);
    return it == attrs.end() ? dflt : it->second.i;
}
float Node::GetFloat(const std::string& k, float dflt) const {
    auto it = attrs.find(k);
    return it == attrs.end() ? dflt : it->second.f;
}
std::vector<int64_t> Node::GetInts(const std::string& k, std::vector<int64_t> dflt) const {
    auto it = attrs.find(k);
    return it == attrs.end() ? dflt : it->second.ints;
}
std::string Node::GetStr(const std::string& k, const std::string& dflt) const {
    auto it = attrs.find(k);
    return it == attrs.end() ? dflt : it->second.s;
}

Model ParseBytes(const uint8_t* data, size_t size) {
    Model m;
    Reader r{data, data + size};
    bool has_graph = false;
    while (!r.Done()) {
        uint64_t key = r.Varint();
        int fno = (int)(key >> 3), wt = (int)(key & 7);
        switch (fno) {
            case 1: m.ir_version = (int64_t)r.Varint(); break;
            case 2: m.producer = r.Str(); break;
            case 7: m.graph = ParseGraph(r.Sub()); has_graph = true; break;
            case 8: {
                Reader o = r.Sub();
                std::string dom;
                int64_t ver = 0;
                while (!o.Done()) {
                    uint64_t k2 = o.Varint();
                    if ((k2 >> 3) == 1) dom = o.Str();
                    else if ((k2 >> 3) == 2) ver = (int64_t)o.Varint();
                    else o.Skip((int)(k2 & 7));
                }
                if (dom.empty() || dom == "ai.onnx") m.opset = ver;
                break;
            }
            default: r.Skip(wt);
        }
    }
    if (!has_graph) throw std::runtime_error("onnx: file has no graph");
    return m;
}

Model ParseFile(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<uint8_t> buf((size_t)n);
    if (n > 0 && !f.read((char*)buf.data(), n)) throw std::runtime_error("cannot read " + path);
    return ParseBytes(buf.data(), buf.size());
}

}  // namespace onnx
}  // namespace b200
