// Host-side mirror of the reference's pybind11 module `_libPolarDecoder`
// (PolarDecoder/PolarDecoder/_cpp/_libPolarDecoder.cpp:29-50 and py_interface/py_*.cpp): the same 15 class
// names, constructor keyword names and `decode` argument names, implemented on top of the C ABI in
// include/polar_b200.h.  Everything heavy happens behind that ABI on the GPU; this file only converts the
// Python objects the reference drivers pass (nested lists / numpy arrays) into the flattened pd_config.
//
// Extensions over the reference surface (all optional, keyword-only in spirit):
//   * decode() also accepts a (B,N) batch and then returns (B,K) -- the entry used for throughput;
//   * every constructor takes device=<int> (default 0);
//   * LUT_f / LUT_g entries may be given as one [Qa][Qb] / [2][Qa][Qb] table per node.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/polar_b200.h"

namespace py = pybind11;

namespace pb200 {

using IArr = py::array_t<int32_t, py::array::c_style | py::array::forcecast>;
using DArr = py::array_t<double, py::array::c_style | py::array::forcecast>;

[[noreturn]] void raise_status(int rc) {
    std::string msg = pd_last_error();
    if (rc == PD_EINVAL || rc == PD_ERANGE) throw py::value_error(msg);
    throw std::runtime_error(msg);
}

struct Tables {
    std::vector<int32_t> frozen, node_type, crc_loc, lut_pool, f_npos, g_npos, f_qa, f_qb, g_qa, g_qb;
    std::vector<int64_t> f_off, g_off, llr_off;
    std::vector<double> llr_pool, r_f, r_g, bf, bg, rf, rg;
    int llr_levels = 0, nb = 0, nr = 0;
};

std::vector<int32_t> to_ivec(const py::object &o) {
    IArr a = IArr::ensure(o);
    if (!a) throw py::value_error("expected an integer sequence");
    return std::vector<int32_t>(a.data(), a.data() + a.size());
}
std::vector<double> to_dvec(const py::object &o) {
    DArr a = DArr::ensure(o);
    if (!a) throw py::value_error("expected a float sequence");
    return std::vector<double>(a.data(), a.data() + a.size());
}

// LUT_f[node][pos][a][b] / LUT_g[node][pos][u][a][b]  (PD/src/SCLUTDecoder.cpp:57,88)
void flatten_lut(const py::object &lut, int N, bool is_g, Tables &t) {
    py::sequence seq = py::reinterpret_borrow<py::sequence>(lut);
    if ((int)py::len(seq) < N - 1) throw py::value_error(std::string(is_g ? "LUT_g" : "LUT_f") + " needs N-1 node entries");
    auto &off = is_g ? t.g_off : t.f_off;
    auto &npos = is_g ? t.g_npos : t.f_npos;
    auto &qa = is_g ? t.g_qa : t.f_qa;
    auto &qb = is_g ? t.g_qb : t.f_qb;
    const int want = is_g ? 4 : 3;
    for (int p = 0; p < N - 1; ++p) {
        IArr a = IArr::ensure(seq[p]);
        if (!a) throw py::value_error("LUT node " + std::to_string(p) + ": not a regular integer array");
        int nd = (int)a.ndim();
        if (nd != want && nd != want - 1) throw py::value_error("LUT node " + std::to_string(p) + ": unexpected rank");
        if (is_g && a.shape(nd - 3) != 2) throw py::value_error("LUT_g node " + std::to_string(p) + ": u axis must have 2 entries");
        off.push_back((int64_t)t.lut_pool.size());
        npos.push_back(nd == want ? (int32_t)a.shape(0) : 1);
        qa.push_back((int32_t)a.shape(nd - 2));
        qb.push_back((int32_t)a.shape(nd - 1));
        t.lut_pool.insert(t.lut_pool.end(), a.data(), a.data() + a.size());
    }
}

// virtual_channel_llr[level][pos][sym], possibly ragged in the last axis
void flatten_llr(const py::object &llr, int N, Tables &t) {
    py::sequence lv = py::reinterpret_borrow<py::sequence>(llr);
    t.llr_levels = (int)py::len(lv);
    t.llr_off.push_back(0);
    for (int l = 0; l < t.llr_levels; ++l) {
        py::object level = lv[l];
        DArr whole;
        try { whole = DArr::ensure(level); } catch (py::error_already_set &) { whole = DArr(); }
        if (!whole) PyErr_Clear();
        if (whole && whole.ndim() == 2 && whole.shape(0) >= N) {
            int q = (int)whole.shape(1);
            for (int pos = 0; pos < N; ++pos) {
                t.llr_pool.insert(t.llr_pool.end(), whole.data() + (size_t)pos * q, whole.data() + (size_t)(pos + 1) * q);
                t.llr_off.push_back((int64_t)t.llr_pool.size());
            }
            continue;
        }
        py::sequence rows = py::reinterpret_borrow<py::sequence>(level);
        if ((int)py::len(rows) < N) throw py::value_error("virtual_channel_llr level needs N rows");
        for (int pos = 0; pos < N; ++pos) {
            std::vector<double> r = to_dvec(rows[pos]);
            t.llr_pool.insert(t.llr_pool.end(), r.begin(), r.end());
            t.llr_off.push_back((int64_t)t.llr_pool.size());
        }
    }
}

// Large decode() results live in pinned host memory (pd_host_alloc): the device->host copy then lands in the numpy array
// itself instead of going through a staging area.  A few buffers are recycled (pinning is slow); the array's capsule hands
// its buffer back when numpy frees the array.
namespace {
struct PinnedPool {
    std::mutex m;
    std::vector<std::pair<void *, size_t>> free_;
    void *get(size_t n, size_t *cap) {
        {
            std::lock_guard<std::mutex> lk(m);
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].second >= n && free_[i].second <= 2 * n + (1u << 20)) {
                    void *p = free_[i].first;
                    *cap = free_[i].second;
                    free_.erase(free_.begin() + (long)i);
                    return p;
                }
        }
        *cap = n;
        return pd_host_alloc(n);
    }
    void put(void *p, size_t cap) {
        {
            std::lock_guard<std::mutex> lk(m);
            if (free_.size() < 4) { free_.emplace_back(p, cap); return; }
        }
        pd_host_free(p);
    }
};
PinnedPool &pinned_pool() { static PinnedPool *p = new PinnedPool(); return *p; }   // (never destroyed: no CUDA calls at exit)
struct PinnedBlock { void *p; size_t cap; };

py::array_t<uint8_t> result_array(int64_t B, int ko, bool flat) {
    const size_t bytes = (size_t)B * (size_t)ko;
    static const size_t min_bytes = getenv("POLAR_B200_PINNED_RESULT_MIN") ? (size_t)atoll(getenv("POLAR_B200_PINNED_RESULT_MIN")) : (size_t)(4u << 20);
    if (flat || bytes < min_bytes) return flat ? py::array_t<uint8_t>(ko) : py::array_t<uint8_t>({(py::ssize_t)B, (py::ssize_t)ko});
    size_t cap = 0;
    void *p = pinned_pool().get(bytes, &cap);
    if (!p) return py::array_t<uint8_t>({(py::ssize_t)B, (py::ssize_t)ko});
    auto *blk = new PinnedBlock{p, cap};
    py::capsule owner(blk, [](void *v) { auto *b = static_cast<PinnedBlock *>(v); pinned_pool().put(b->p, b->cap); delete b; });
    return py::array_t<uint8_t>({(py::ssize_t)B, (py::ssize_t)ko}, {(py::ssize_t)ko, (py::ssize_t)1}, static_cast<uint8_t *>(p), owner);
}
}  // namespace

class Decoder {
public:
    Decoder() = default;
    Decoder(const Decoder &) = delete;
    Decoder &operator=(const Decoder &) = delete;
    ~Decoder() { pd_destroy(dec_); }

    void init(int kind, int N, int K, int A, int L, const py::object &frozen, const py::object &node_type, int crc_n,
              const py::object &crc_p, const py::object &lut_f, const py::object &lut_g, const py::object &llr,
              const py::object &r_f, const py::object &r_g, int v, const py::object &bf, const py::object &bg,
              const py::object &rf, const py::object &rg, int device) {
        Tables t;
        pd_config c{};
        c.kind = kind; c.N = N; c.K = K; c.A = A; c.L = L; c.device = device; c.v = v;
        t.frozen = to_ivec(frozen);
        if ((int)t.frozen.size() < N) throw py::value_error("frozen_bits must have N entries");
        c.frozen_bits = t.frozen.data();
        if (!node_type.is_none()) {
            t.node_type = to_ivec(node_type);
            if ((int)t.node_type.size() < N - 1) throw py::value_error("node_type must cover all internal nodes");
            t.node_type.resize(2 * (size_t)N - 1, -1);
            c.node_type = t.node_type.data();
        }
        if (!crc_p.is_none()) {
            t.crc_loc = to_ivec(crc_p);
            c.crc_n = crc_n; c.crc_loc = t.crc_loc.data(); c.crc_loc_len = (int)t.crc_loc.size();
        }
        if (!lut_f.is_none()) {
            flatten_lut(lut_f, N, false, t);
            flatten_lut(lut_g, N, true, t);
            flatten_llr(llr, N, t);
            c.lut_pool = t.lut_pool.data(); c.lut_pool_len = (int64_t)t.lut_pool.size();
            c.f_off = t.f_off.data(); c.g_off = t.g_off.data();
            c.f_npos = t.f_npos.data(); c.g_npos = t.g_npos.data();
            c.f_qa = t.f_qa.data(); c.f_qb = t.f_qb.data(); c.g_qa = t.g_qa.data(); c.g_qb = t.g_qb.data();
            c.llr_pool = t.llr_pool.data(); c.llr_off = t.llr_off.data(); c.llr_levels = t.llr_levels;
        }
        if (!r_f.is_none()) {
            t.r_f = to_dvec(r_f); t.r_g = to_dvec(r_g);
            if ((int)t.r_f.size() < N - 1 || (int)t.r_g.size() < N - 1) throw py::value_error("decoder_r_f / decoder_r_g need N-1 entries");
            c.decoder_r_f = t.r_f.data(); c.decoder_r_g = t.r_g.data();
        }
        if (!bf.is_none()) {
            auto grab = [&](const py::object &o, std::vector<double> &dst, int &width) {
                DArr a = DArr::ensure(o);
                if (!a || a.ndim() != 2 || a.shape(0) < N - 1) throw py::value_error("Lloyd tables must be [N-1][width] arrays");
                width = (int)a.shape(1);
                dst.assign(a.data(), a.data() + (size_t)(N - 1) * width);
            };
            int wb = 0, wb2 = 0, wr = 0, wr2 = 0;
            grab(bf, t.bf, wb); grab(bg, t.bg, wb2); grab(rf, t.rf, wr); grab(rg, t.rg, wr2);
            if (wb != wb2 || wr != wr2) throw py::value_error("f and g Lloyd tables must have equal widths");
            c.boundaries_f = t.bf.data(); c.boundaries_g = t.bg.data();
            c.reconstruction_f = t.rf.data(); c.reconstruction_g = t.rg.data();
            c.n_boundaries = wb; c.n_reconstruction = wr;
        }
        int rc = pd_create(&c, &dec_);
        if (rc != PD_OK) raise_status(rc);
        N_ = N;
        lut_ = !lut_f.is_none();
    }

    // decode((N,)) / ((1,N)) -> (K,) like the reference; decode((B,N)) -> (B,K)
    py::array_t<uint8_t> decode(const py::object &x) {
        const int ko = pd_out_len(dec_);
        py::array arr;
        int dtype;
        if (lut_) {
            py::array in = py::array::ensure(x);
            if (!in) throw py::value_error("decode: expected an array");
            if (py::isinstance<py::array_t<uint8_t>>(in) && (in.flags() & py::array::c_style)) {
                arr = in; dtype = PD_U8;
            } else if (py::isinstance<py::array_t<double>>(in) && (in.flags() & py::array::c_style)) {
                arr = in; dtype = PD_F64;   // float64-typed symbols (probability-domain driver): pd_decode truncates them like the cast below
            } else {
                arr = IArr::ensure(x);   // the reference takes py::array_t<int> with forcecast
                dtype = PD_I32;
            }
        } else {
            arr = DArr::ensure(x);
            dtype = PD_F64;
        }
        if (!arr) throw py::value_error("decode: cannot convert the input");
        int64_t B;
        bool flat;
        if (arr.ndim() == 1) {
            if (arr.shape(0) < N_) throw py::value_error("decode: need at least N values");
            B = 1; flat = true;   // the reference reads the first N elements only
        } else if (arr.ndim() == 2 && arr.shape(1) == N_) {
            B = arr.shape(0); flat = (B == 1);
        } else {
            throw py::value_error("decode: expected shape (N,), (1,N) or (B,N)");
        }
        py::array_t<uint8_t> out = result_array(B, ko, flat);
        int rc;
        {
            py::gil_scoped_release nogil;
            std::lock_guard<std::mutex> one_call_at_a_time(mu_);   // the C ABI object is not re-entrant; the reference (GIL held) was
            rc = pd_decode(dec_, arr.data(), dtype, B, out.mutable_data());
        }
        if (rc != PD_OK) raise_status(rc);
        return out;
    }

    // shape handling shared by the blind-detection entries: (N,), (1,N) -> one frame; (B,N) -> batch
    DArr llr_frames(const py::object &x, int64_t &B, bool &flat) const {
        DArr arr = DArr::ensure(x);
        if (!arr) throw py::value_error("expected a float array");
        if (arr.ndim() == 1) {
            if (arr.shape(0) < N_) throw py::value_error("need at least N values");
            B = 1; flat = true;
        } else if (arr.ndim() == 2 && arr.shape(1) == N_) {
            B = arr.shape(0); flat = (B == 1);
        } else {
            throw py::value_error("expected shape (N,), (1,N) or (B,N)");
        }
        return arr;
    }
    // DMetricCalculator.calculate (PolarEncoder/PolarBD/_cpp/src/DMetric.cpp:25): float for one frame, (B,) for a batch
    py::object bd_calculate(const py::object &x) {
        int64_t B; bool flat;
        DArr arr = llr_frames(x, B, flat);
        py::array_t<double> metric((py::ssize_t)B);
        int rc;
        {
            py::gil_scoped_release nogil;
            std::lock_guard<std::mutex> one_call_at_a_time(mu_);   // the C ABI object is not re-entrant; the reference (GIL held) was
            rc = pd_decode_bd(dec_, arr.data(), B, nullptr, 0, nullptr, metric.mutable_data(), nullptr);
        }
        if (rc != PD_OK) raise_status(rc);
        if (flat) return py::float_(metric.data()[0]);
        return std::move(metric);
    }
    // PolarBD CASCL::decode(llr, RNTI) -> (bits, PM, isPass)  (CASCLWithRNTI.cpp:74,251)
    py::tuple bd_decode(const py::object &x, const py::object &rnti) {
        int64_t B; bool flat;
        DArr arr = llr_frames(x, B, flat);
        std::vector<int32_t> r = to_ivec(rnti);
        const int ko = pd_out_len(dec_);
        py::array_t<uint8_t> bits = flat ? py::array_t<uint8_t>(ko) : py::array_t<uint8_t>({(py::ssize_t)B, (py::ssize_t)ko});
        py::array_t<double> pm((py::ssize_t)B);
        py::array_t<uint8_t> pass((py::ssize_t)B);
        int rc;
        {
            py::gil_scoped_release nogil;
            std::lock_guard<std::mutex> one_call_at_a_time(mu_);   // the C ABI object is not re-entrant; the reference (GIL held) was
            rc = pd_decode_bd(dec_, arr.data(), B, r.data(), (int32_t)r.size(), bits.mutable_data(), pm.mutable_data(), pass.mutable_data());
        }
        if (rc != PD_OK) raise_status(rc);
        if (flat) return py::make_tuple(bits, py::float_(pm.data()[0]), py::bool_(pass.data()[0] != 0));
        return py::make_tuple(bits, pm, pass.attr("astype")("bool"));
    }

    // decode() deals its batch to these CUDA devices from now on (pd_set_devices); [] = the constructor's device alone
    void set_devices(const std::vector<int32_t> &ids) {
        std::lock_guard<std::mutex> one_call_at_a_time(mu_);
        const int rc = pd_set_devices(dec_, (int32_t)ids.size(), ids.data());
        if (rc != PD_OK) raise_status(rc);
    }
    int device_count() const { return pd_device_count(dec_); }
    std::string kernel() const { return pd_kernel_name(dec_); }
    std::string kernel_note() const { return pd_kernel_note(dec_); }
    uintptr_t handle() const { return reinterpret_cast<uintptr_t>(dec_); }

private:
    pd_decoder *dec_ = nullptr;
    std::mutex mu_;
    int N_ = 0;
    bool lut_ = false;
};

#define NONE py::none()

// One distinct C++ type per reference class so that pybind11 registers 15 independent Python classes.
template <int KIND> struct Cls : Decoder {};

template <int KIND, typename... Extra>
py::class_<Cls<KIND>> declare(py::module_ &m, const char *name, const char *doc, const char *decode_arg) {
    py::class_<Cls<KIND>> c(m, name, doc);
    c.def("decode", &Decoder::decode, py::arg(decode_arg));
    c.def_property_readonly("kernel", &Decoder::kernel, "name of the CUDA kernel variant in use");
    c.def_property_readonly("kernel_note", &Decoder::kernel_note, "why a LUT decoder is not on scl_lut_warp ('' when it is)");
    c.def("set_devices", &Decoder::set_devices, py::arg("device_ids"), "shard every decode() call over these CUDA devices of the box (pd_set_devices)");
    c.def_property_readonly("device_count", &Decoder::device_count);
    c.def_property_readonly("_handle", &Decoder::handle, "pd_decoder* for direct C-ABI calls");
    return c;
}

}  // namespace pb200

using namespace pb200;
using py::arg;

PYBIND11_MODULE(_libPolarDecoder, m) {
    m.doc() = "Decoders For Polar Codes (B200 / sm_100a build behind the reference PolarDecoder API)";
    m.attr("backend") = "polar_b200";
    m.def("version", []() { return std::string(pd_version()); });
    m.def("launch_count", []() { return pd_launch_count(); });

#define MK(KIND) std::unique_ptr<Cls<KIND>> self(new Cls<KIND>())

    // PD/py_interface/py_SCDecoder.cpp:9-11
    declare<PD_SC>(m, "SCDecoder", "Successive Cancellation Decoder", "llr")
        .def(py::init([](int N, int K, py::object fb, py::object mb, int device) {
                 MK(PD_SC); self->init(PD_SC, N, K, 0, 1, fb, NONE, 0, NONE, NONE, NONE, NONE, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("frozen_bits"), arg("message_bits"), arg("device") = 0);
    // py_FastSCDecoder.cpp:10-12
    declare<PD_FASTSC>(m, "FastSCDecoder", "Fast Successive Cancellation Decoder", "llr")
        .def(py::init([](int N, int K, py::object fb, py::object mb, py::object nt, int device) {
                 MK(PD_FASTSC); self->init(PD_FASTSC, N, K, 0, 1, fb, nt, 0, NONE, NONE, NONE, NONE, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("frozen_bits"), arg("message_bits"), arg("node_type"), arg("device") = 0);
    // py_SCLDecoder.cpp:10-12
    declare<PD_SCL>(m, "SCLDecoder", "Successive Cancellation List Decoder", "llr")
        .def(py::init([](int N, int K, int L, py::object fb, py::object mb, int device) {
                 MK(PD_SCL); self->init(PD_SCL, N, K, 0, L, fb, NONE, 0, NONE, NONE, NONE, NONE, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("device") = 0);
    // py_FastSCLDecoder.cpp:9-11
    declare<PD_FASTSCL>(m, "FastSCLDecoder", "Fast Successive Cancellation List Decoder", "llr")
        .def(py::init([](int N, int K, int L, py::object fb, py::object mb, py::object nt, int device) {
                 MK(PD_FASTSCL); self->init(PD_FASTSCL, N, K, 0, L, fb, nt, 0, NONE, NONE, NONE, NONE, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("node_type"), arg("device") = 0);
    // py_CASCLDecoder.cpp:9-11
    declare<PD_CASCL>(m, "CASCLDecoder", "CRC Aided Successive Cancellation List Decoder", "llr")
        .def(py::init([](int N, int K, int A, int L, py::object fb, py::object mb, int crc_n, py::object crc_p, int device) {
                 MK(PD_CASCL); self->init(PD_CASCL, N, K, A, L, fb, NONE, crc_n, crc_p, NONE, NONE, NONE, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("A"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("crc_n"), arg("crc_p"), arg("device") = 0);
    // py_SCLUTDecoder.cpp:10-14
    declare<PD_SCLUT>(m, "SCLUTDecoder", "Successive Cancellation Decoder Using LUT", "channel_quantized_symbols")
        .def(py::init([](int N, int K, py::object fb, py::object mb, py::object f, py::object g, py::object llr, int device) {
                 MK(PD_SCLUT); self->init(PD_SCLUT, N, K, 0, 1, fb, NONE, 0, NONE, f, g, llr, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("frozen_bits"), arg("message_bits"), arg("LUT_f"), arg("LUT_g"), arg("virtual_channel_llr"), arg("device") = 0);
    // py_FastSCLUTDecoder.cpp:11-15 (note LUT_Fs / LUT_Gs and decode(llr))
    declare<PD_FASTSCLUT>(m, "FastSCLUTDecoder", "Fast Successive Cancellation Decoder Using LUT", "llr")
        .def(py::init([](int N, int K, py::object fb, py::object mb, py::object nt, py::object f, py::object g, py::object llr, int device) {
                 MK(PD_FASTSCLUT); self->init(PD_FASTSCLUT, N, K, 0, 1, fb, nt, 0, NONE, f, g, llr, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("frozen_bits"), arg("message_bits"), arg("node_type"), arg("LUT_Fs"), arg("LUT_Gs"), arg("virtual_channel_llr"), arg("device") = 0);
    // py_SCLLUTDecoder.cpp:11-15
    declare<PD_SCLLUT>(m, "SCLLUTDecoder", "Successive Cancellation List Decoder Using LUT", "channel_quantized_symbols")
        .def(py::init([](int N, int K, int L, py::object fb, py::object mb, py::object f, py::object g, py::object llr, int device) {
                 MK(PD_SCLLUT); self->init(PD_SCLLUT, N, K, 0, L, fb, NONE, 0, NONE, f, g, llr, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("LUT_f"), arg("LUT_g"), arg("virtual_channel_llr"), arg("device") = 0);
    // py_FastSCLLUTDecoder.cpp:12-16
    declare<PD_FASTSCLLUT>(m, "FastSCLLUTDecoder", "Fast Successive Cancellation List Decoder Using LUT", "channel_quantized_symbols")
        .def(py::init([](int N, int K, int L, py::object fb, py::object mb, py::object nt, py::object f, py::object g, py::object llr, int device) {
                 MK(PD_FASTSCLLUT); self->init(PD_FASTSCLLUT, N, K, 0, L, fb, nt, 0, NONE, f, g, llr, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("node_type"), arg("LUT_f"), arg("LUT_g"), arg("virtual_channel_llr"), arg("device") = 0);
    // py_CASCLLUTDecoder.cpp:10-16
    declare<PD_CASCLLUT>(m, "CASCLLUTDecoder", "CRC Aided Successive Cancellation List Decoder Using LUT", "channel_quantized_symbols")
        .def(py::init([](int N, int K, int A, int L, py::object fb, py::object mb, int crc_n, py::object crc_p, py::object f, py::object g, py::object llr, int device) {
                 MK(PD_CASCLLUT); self->init(PD_CASCLLUT, N, K, A, L, fb, NONE, crc_n, crc_p, f, g, llr, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("A"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("crc_n"), arg("crc_p"), arg("LUT_f"), arg("LUT_g"), arg("virtual_channel_llr"), arg("device") = 0);
    // py_CAFastSCLLUTDecoder.cpp:11-15
    declare<PD_CAFASTSCLLUT>(m, "CAFastSCLLUTDecoder", "CRC Aided Fast Successive Cancellation List Decoder Using LUT", "channel_quantized_symbols")
        .def(py::init([](int N, int K, int A, int L, py::object fb, py::object mb, py::object nt, py::object f, py::object g, py::object llr, int device) {
                 MK(PD_CAFASTSCLLUT); self->init(PD_CAFASTSCLLUT, N, K, A, L, fb, nt, 0, NONE, f, g, llr, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("A"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("node_type"), arg("LUT_f"), arg("LUT_g"), arg("virtual_channel_llr"), arg("device") = 0);
    // py_SCUniformDecoder.cpp:10-15
    declare<PD_SC_UNIFORM>(m, "SCUniformQuantizedDecoder", "Uniformly Quantized Successive Cancellation Decoder", "llr")
        .def(py::init([](int N, int K, py::object fb, py::object mb, py::object rf, py::object rg, int v, int device) {
                 MK(PD_SC_UNIFORM); self->init(PD_SC_UNIFORM, N, K, 0, 1, fb, NONE, 0, NONE, NONE, NONE, NONE, rf, rg, v, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("frozen_bits"), arg("message_bits"), arg("decoder_r_f"), arg("decoder_r_g"), arg("v"), arg("device") = 0);
    // py_SCLUniformQuantizedDecoder.cpp:11-16
    declare<PD_SCL_UNIFORM>(m, "SCLUniformQuantizedDecoder", "Uniformly Quantized Successive Cancellation List Decoder", "llr")
        .def(py::init([](int N, int K, int L, py::object fb, py::object mb, py::object rf, py::object rg, int v, int device) {
                 MK(PD_SCL_UNIFORM); self->init(PD_SCL_UNIFORM, N, K, 0, L, fb, NONE, 0, NONE, NONE, NONE, NONE, rf, rg, v, NONE, NONE, NONE, NONE, device); return self; }),
             arg("N"), arg("K"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("decoder_r_f"), arg("decoder_r_g"), arg("v"), arg("device") = 0);
    // py_SCLloydQuantizedDecoder.cpp:11-18
    declare<PD_SC_LLOYD>(m, "SCLloydQuantizedDecoder", "Lloyd Quantized Successive Cancellation Decoder", "llr")
        .def(py::init([](int N, int K, py::object fb, py::object mb, py::object bf, py::object bg, py::object rf, py::object rg, int v, int device) {
                 MK(PD_SC_LLOYD); self->init(PD_SC_LLOYD, N, K, 0, 1, fb, NONE, 0, NONE, NONE, NONE, NONE, NONE, NONE, v, bf, bg, rf, rg, device); return self; }),
             arg("N"), arg("K"), arg("frozen_bits"), arg("message_bits"), arg("boundaries_f"), arg("boundaries_g"), arg("reconstruction_f"), arg("reconstruction_g"), arg("v"), arg("device") = 0);
    // py_SCLLloydQuantizedDecoder.cpp:11-18
    declare<PD_SCL_LLOYD>(m, "SCLLloydQuantizedDecoder", "Lloyd Quantized Successive Cancellation List Decoder", "llr")
        .def(py::init([](int N, int K, int L, py::object fb, py::object mb, py::object bf, py::object bg, py::object rf, py::object rg, int v, int device) {
                 MK(PD_SCL_LLOYD); self->init(PD_SCL_LLOYD, N, K, 0, L, fb, NONE, 0, NONE, NONE, NONE, NONE, NONE, NONE, v, bf, bg, rf, rg, device); return self; }),
             arg("N"), arg("K"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("boundaries_f"), arg("boundaries_g"), arg("reconstruction_f"), arg("reconstruction_g"), arg("v"), arg("device") = 0);
    // Blind-detection helpers of PolarEncoder/PolarBD (module libPolarBD there; re-exported under the reference's names by
    // quantized_decoder_polar_codes_b200/PolarBD):  py_DMetric.cpp:10-13, py_CASCLWithRNTI.cpp:8-11
    {
        py::class_<Cls<PD_BD_DMETRIC>> c(m, "BDDMetricCalculator", "DMetric Calculator with Fast-SSC");
        c.def(py::init([](int N, int K, py::object fb, py::object mb, py::object nt, int device) {
                  MK(PD_BD_DMETRIC); self->init(PD_BD_DMETRIC, N, K, 0, 1, fb, nt, 0, NONE, NONE, NONE, NONE, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
              arg("N"), arg("K"), arg("frozen_bits"), arg("message_bits"), arg("node_type"), arg("device") = 0);
        c.def("calculate", &Decoder::bd_calculate, arg("llr"));
        c.def_property_readonly("kernel", &Decoder::kernel);
        c.def_property_readonly("_handle", &Decoder::handle);
    }
    {
        py::class_<Cls<PD_BD_CASCL>> c(m, "BDCASCLDecoder", "CRC Aided Successive Cancellation List Decoder (RNTI-scrambled CRC)");
        c.def(py::init([](int N, int K, int A, int L, py::object fb, py::object mb, int crc_n, py::object crc_p, int device) {
                  MK(PD_BD_CASCL); self->init(PD_BD_CASCL, N, K, A, L, fb, NONE, crc_n, crc_p, NONE, NONE, NONE, NONE, NONE, 0, NONE, NONE, NONE, NONE, device); return self; }),
              arg("N"), arg("K"), arg("A"), arg("L"), arg("frozen_bits"), arg("message_bits"), arg("crc_n"), arg("crc_p"), arg("device") = 0);
        c.def("decode", &Decoder::bd_decode, arg("llr"), arg("RNTI"));
        c.def_property_readonly("kernel", &Decoder::kernel);
        c.def_property_readonly("_handle", &Decoder::handle);
    }
#undef MK
}
