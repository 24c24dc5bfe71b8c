// groan_gpu.hpp -- C++17 host-side mirror of the reference's System / Group / SimBox / Dimension interface for the
// per-frame PBC geometry path, header-only, on top of the C ABI of groan_gpu.h (libgroan_gpu.so).
//
// groan_rs is Rust and this image has no Rust toolchain, so the host side above the C ABI is written in C++ with the
// reference's names, argument meaning and error behaviour (INTEGRATION.md shows the Rust binding a maintainer would add
// for the same entry points).  Every method evaluates the reference function it is named after for EVERY frame of the
// current batch and returns one result per frame.  There is no CPU implementation behind any of them.
//
//   reference                                            here
//   System::group_get_center (analysis.rs:105)           System::group_get_center(name)  -> std::vector<Vector3D>, one per frame
//   System::calc_rmsd (rmsd.rs:75)                       System::calc_rmsd(reference, name) -> std::vector<float>
//   GroupError::NotFound / EmptyGroup / InvalidSimBox    exceptions GroupError{variant = "NotFound" / ...}
//   traj_iter_map_reduce body (parallel.rs:208-269)      FrameBatcher: buffers frames, flushes a batch to the GPU every B frames
//   XtcReader / GroupXtcReader (xtc_io/mod.rs, molly_xtc.rs)  XtcFile + System::set_frames_xtc / set_group_frames / write_xtc
//   System::guess_bonds (guess.rs:362)                   System::guess_bonds(vdw, factor) -> bonds per frame
//   HBondAnalysis (hbonds.rs:160-335)                    HBondAnalysis{chains, pairs, max_distance, min_angle}.analyze(system)
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <tuple>
#include <string>
#include <utility>
#include <vector>

#include "groan_gpu.h"
#include "groan_xtc.h"

namespace groan {

// src/structures/dimension.rs:15-25
enum class Dimension : int { None = 0, X = 1, Y = 2, Z = 3, XY = 4, XZ = 5, YZ = 6, XYZ = 7 };

// src/structures/vector3d.rs (only as a value type: all arithmetic happens on the device)
using Vector3D = std::array<float, 3>;

// src/structures/simbox.rs:13-26.  Stored as the row-major 3x3 matrix the xtc reader emits (io/xdrfile.rs:170-187).
struct SimBox {
    float m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    static SimBox orthogonal(float x, float y, float z) {
        SimBox b;
        b.m[0] = x; b.m[4] = y; b.m[8] = z;
        return b;
    }
    // From<[f32; 9]>, simbox.rs:28-52: v1x v2y v3z v1y v1z v2x v2z v3x v3y (the order of a .gro box line)
    static SimBox from_gro(const std::array<float, 9> &g) {
        SimBox b;
        b.m[0] = g[0]; b.m[1] = g[3]; b.m[2] = g[4];
        b.m[3] = g[5]; b.m[4] = g[1]; b.m[5] = g[6];
        b.m[6] = g[7]; b.m[7] = g[8]; b.m[8] = g[2];
        return b;
    }
    bool is_orthogonal() const { return m[1] == 0 && m[2] == 0 && m[3] == 0 && m[5] == 0 && m[6] == 0 && m[7] == 0; } // simbox.rs:185
};

// ---- errors (src/errors.rs): one exception type per reference enum, `variant` = the enum variant ----------------
struct GroanError : std::runtime_error {
    int status;
    std::string variant;
    size_t a = 0, b = 0; // PositionError / MassError: (frame, atom); RMSDError::InconsistentGroup: (n_ref, n_target)
    GroanError(const std::string &v, const std::string &msg, int st) : std::runtime_error(v + ": " + msg), status(st), variant(v) {}
};
struct SimBoxError : GroanError { using GroanError::GroanError; };   // errors.rs:556-582
struct GroupError : GroanError { using GroanError::GroanError; };    // errors.rs:624-650
struct RMSDError : GroanError { using GroanError::GroanError; };     // rmsd.rs errors
struct GpuError : GroanError { using GroanError::GroanError; };      // CUDA failure / misuse of the batch interface

class System {
  public:
    // n_atoms atoms; up to max_frames frames per batch; one System = one GPU context = one host thread
    explicit System(size_t n_atoms, size_t max_frames = 1, int device = 0) : n_atoms_(n_atoms), max_frames_(max_frames) {
        const int st = groan_gpu_create(device, n_atoms, max_frames, &ctx_);
        if (st != GROAN_OK) throw GpuError("Create", groan_gpu_strerror(st), st);
        groups_["all"] = Group{GROAN_GROUP_ALL, {}};
    }
    ~System() {
        if (ctx_) groan_gpu_destroy(ctx_);
    }
    System(const System &) = delete;
    System &operator=(const System &) = delete;

    size_t get_n_atoms() const { return n_atoms_; }
    size_t n_frames() const { return n_frames_; }

    // masses per atom (Atom::mass is Option<f32>: a negative value = "no mass"); apply before creating groups
    void set_masses(std::vector<float> masses) {
        if (masses.size() != n_atoms_) throw GpuError("InvalidArgument", "one mass per atom expected", GROAN_EINVAL);
        masses_ = std::move(masses);
        // "all" exists from System::new on with the atoms' masses (src/system/mod.rs:150-170)
        check(groan_gpu_set_group(ctx_, GROAN_GROUP_ALL, nullptr, n_atoms_, masses_.data()), "set_masses");
        for (auto &kv : groups_)
            if (kv.second.gid >= 0) upload_group(kv.second);
        ref_keys_.clear();
        version_++;
    }

    // System::group_create_from_indices (groups.rs) -> Group::from_indices (container.rs:51-115): sorted, de-duplicated
    void group_create_from_indices(const std::string &name, std::vector<uint32_t> indices) {
        std::sort(indices.begin(), indices.end());
        indices.erase(std::unique(indices.begin(), indices.end()), indices.end());
        if (!indices.empty() && indices.back() >= n_atoms_) throw GroupError("InvalidIndex", name, GROAN_EINVAL);
        auto it = groups_.find(name);
        Group g;
        if (it != groups_.end() && it->second.gid >= 0) g.gid = it->second.gid; // overwrite, like the reference (with a warning there)
        else if (next_gid_ >= GROAN_MAX_GROUPS) throw GpuError("Capacity", "too many groups", GROAN_ECAPACITY);
        else g.gid = next_gid_++;
        g.indices = std::move(indices);
        upload_group(g);
        groups_[name] = std::move(g);
        ref_keys_.erase(name); // groan_gpu_set_group dropped the device-side RMSD reference of this group id
        version_++;            // a System that serves as somebody's reference has changed
    }
    void group_create_from_range(const std::string &name, uint32_t first, uint32_t last_inclusive) {
        std::vector<uint32_t> idx;
        for (uint32_t i = first; i <= last_inclusive; i++) idx.push_back(i);
        group_create_from_indices(name, std::move(idx));
    }
    bool group_exists(const std::string &name) const { return groups_.count(name) != 0; }
    size_t group_get_n_atoms(const std::string &name) const {
        const Group &g = group(name, false);
        return g.gid == GROAN_GROUP_ALL ? n_atoms_ : g.indices.size();
    }

    // ---- frames: what FrameData::update_system would copy into Vec<Atom> (traj_read.rs:50-63), for a whole batch
    // xyz: n_frames x n_atoms x 3; boxes: n_frames entries or nullptr (= the system has no box)
    void set_frames(const float *xyz, const SimBox *boxes, size_t n_frames) {
        std::vector<float> b;
        if (boxes) {
            b.resize(n_frames * 9);
            for (size_t f = 0; f < n_frames; f++) std::memcpy(&b[f * 9], boxes[f].m, sizeof(boxes[f].m));
        }
        check(groan_gpu_push_frames(ctx_, xyz, boxes ? b.data() : nullptr, n_frames), "set_frames");
        n_frames_ = n_frames;
        frame0_.assign(xyz, xyz + n_atoms_ * 3); // a System used as an RMSD reference is read on the host (rmsd.rs:186-203)
        frame0_stale_ = false;
        has_box0_ = boxes != nullptr;
        if (boxes) box0_ = boxes[0];
        version_++;
    }
    void set_frame(const std::vector<Vector3D> &positions, const SimBox &box) {
        if (positions.size() != n_atoms_) throw GpuError("InvalidArgument", "one position per atom expected", GROAN_EINVAL);
        set_frames(&positions[0][0], &box, 1);
    }
    std::vector<float> get_frames() {
        std::vector<float> out(n_frames_ * n_atoms_ * 3);
        check(groan_gpu_get_frames(ctx_, out.data()), "get_frames");
        return out;
    }
    void sync() { check(groan_gpu_sync(ctx_), "sync"); }

    // ---- centres (src/system/analysis.rs:52,105,258; iterators.rs)
    std::vector<Vector3D> group_estimate_center(const std::string &name) { return centre(name, 0, 0); }
    std::vector<Vector3D> group_estimate_com(const std::string &name) { return centre(name, 0, 1); }
    std::vector<Vector3D> group_get_center(const std::string &name) { return centre(name, 1, 0); }
    std::vector<Vector3D> group_get_com(const std::string &name) { return centre(name, 1, 1); }
    std::vector<Vector3D> group_get_center_naive(const std::string &name) { return centre(name, 2, 0); }

    // ---- distances (analysis.rs:348-427)
    std::vector<float> group_distance(const std::string &g1, const std::string &g2, Dimension dim) {
        std::vector<float> out(n_frames_);
        check(groan_gpu_group_distance(ctx_, group(g1).gid, group(g2).gid, (int)dim, out.data()), "group_distance", g1);
        return out;
    }
    // System::atoms_distance (analysis.rs:459-471), per frame
    std::vector<float> atoms_distance(uint32_t index1, uint32_t index2, Dimension dim) {
        group_create_from_indices("__atom1", {index1});
        group_create_from_indices("__atom2", {index2});
        return group_all_distances("__atom1", "__atom2", dim);
    }
    // n_frames x n1 x n2, row-major like ndarray::Array2 per frame
    std::vector<float> group_all_distances(const std::string &g1, const std::string &g2, Dimension dim) {
        std::vector<float> out(n_frames_ * group_get_n_atoms(g1) * group_get_n_atoms(g2));
        check(groan_gpu_all_distances(ctx_, group(g1).gid, group(g2).gid, (int)dim, out.data()), "group_all_distances", g1);
        return out;
    }

    // ---- cutoff pair search: CellGrid::new(g2) + neighbors_iter around every atom of g1 + distance filter (cellgrid.rs:301-420)
    struct PairList {
        std::vector<uint64_t> count;                        // per frame, always complete
        std::vector<std::array<uint32_t, 2>> pairs;         // n_frames x capacity (positions inside g1, g2), order undefined
        std::vector<float> dist;                            // n_frames x capacity
        size_t capacity = 0;
    };
    PairList group_pairs_within(const std::string &g1, const std::string &g2, float cutoff, size_t capacity = 0) {
        PairList out;
        out.capacity = capacity;
        out.count.assign(n_frames_, 0);
        out.pairs.resize(n_frames_ * capacity);
        out.dist.resize(n_frames_ * capacity);
        check(groan_gpu_pairs_within(ctx_, group(g1).gid, group(g2).gid, cutoff, out.count.data(),
                                     capacity ? &out.pairs[0][0] : nullptr, capacity ? out.dist.data() : nullptr, capacity),
              "group_pairs_within", g1);
        return out;
    }

    // ---- modifying (src/system/modifying.rs, utility.rs): in place on every frame of the batch
    void atoms_wrap() { check(groan_gpu_wrap(ctx_, GROAN_GROUP_ALL, nullptr), "atoms_wrap"); touched(); }
    void group_wrap(const std::string &name) { check(groan_gpu_wrap(ctx_, group(name).gid, nullptr), "group_wrap", name); touched(); }
    void atoms_translate(const Vector3D &t) { check(groan_gpu_translate(ctx_, GROAN_GROUP_ALL, t.data(), nullptr), "atoms_translate"); touched(); }
    void group_translate(const std::string &name, const Vector3D &t) {
        check(groan_gpu_translate(ctx_, group(name).gid, t.data(), nullptr), "group_translate", name);
        touched();
    }
    void make_group_whole(const std::string &name) {
        check(groan_gpu_make_group_whole(ctx_, group(name).gid), "make_group_whole", name);
        touched();
    }
    // bonds as index pairs (System::add_bonds_from_pdb); molecules and their reference atoms are worked out here, on the host
    void add_bonds(const std::vector<std::pair<uint32_t, uint32_t>> &bonds) {
        std::vector<uint32_t> parent(n_atoms_);
        for (size_t i = 0; i < n_atoms_; i++) parent[i] = (uint32_t)i;
        std::function<uint32_t(uint32_t)> find = [&](uint32_t a) {
            while (parent[a] != a) a = parent[a] = parent[parent[a]];
            return a;
        };
        std::vector<char> bonded(n_atoms_, 0);
        for (const auto &bd : bonds) {
            if (bd.first >= n_atoms_ || bd.second >= n_atoms_) throw GpuError("InvalidArgument", "bond index out of range", GROAN_EINVAL);
            bonded[bd.first] = bonded[bd.second] = 1;
            const uint32_t ra = find(bd.first), rb = find(bd.second);
            if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
        }
        mol_ref_.assign(n_atoms_, GROAN_NO_MOLECULE);
        for (size_t i = 0; i < n_atoms_; i++)
            if (bonded[i]) mol_ref_[i] = find((uint32_t)i); // System::create_mol_references, modifying.rs:258-283
        check(groan_gpu_set_molecules(ctx_, mol_ref_.data()), "add_bonds");
        for (const auto &bd : bonds) bonds_.emplace_back(std::min(bd.first, bd.second), std::max(bd.first, bd.second));
        std::sort(bonds_.begin(), bonds_.end());
        bonds_.erase(std::unique(bonds_.begin(), bonds_.end()), bonds_.end());
    }
    void make_molecules_whole() {
        if (mol_ref_.empty()) {
            mol_ref_.assign(n_atoms_, GROAN_NO_MOLECULE);
            check(groan_gpu_set_molecules(ctx_, mol_ref_.data()), "make_molecules_whole");
        }
        check(groan_gpu_make_molecules_whole(ctx_), "make_molecules_whole");
        touched();
    }
    void atoms_center(const std::string &reference, Dimension dim) {
        check(groan_gpu_atoms_center(ctx_, group(reference).gid, 0, (int)dim), "atoms_center", reference);
        touched();
    }
    void atoms_center_mass(const std::string &reference, Dimension dim) {
        check(groan_gpu_atoms_center(ctx_, group(reference).gid, 1, (int)dim), "atoms_center_mass", reference);
        touched();
    }

    // ---- RMSD (src/system/rmsd.rs:75,129).  `reference` may be another System with its own atom count; the group is looked
    // up by name in both (rmsd.rs:823-841); masses come from the reference system (rmsd.rs:154,192)
    std::vector<float> calc_rmsd(const System &reference, const std::string &name) {
        set_reference(reference, name);
        std::vector<float> out(n_frames_);
        check(groan_gpu_rmsd(ctx_, group(name, true).gid, out.data(), nullptr), "calc_rmsd", name, true);
        return out;
    }
    std::vector<float> calc_rmsd_and_fit(const System &reference, const std::string &name) {
        set_reference(reference, name);
        std::vector<float> out(n_frames_);
        check(groan_gpu_rmsd_fit(ctx_, group(name, true).gid, out.data()), "calc_rmsd_and_fit", name, true);
        touched();
        return out;
    }
    // group_get_center (or group_get_com) AND calc_rmsd from one read of every frame
    std::pair<std::vector<Vector3D>, std::vector<float>> group_center_and_rmsd(const System &reference, const std::string &name,
                                                                              bool weighted = false) {
        set_reference(reference, name);
        std::vector<Vector3D> c(n_frames_);
        std::vector<float> r(n_frames_);
        check(groan_gpu_center_rmsd(ctx_, group(name, true).gid, weighted ? 1 : 0, &c[0][0], r.data(), nullptr), "group_center_and_rmsd",
              name, true);
        return {std::move(c), std::move(r)};
    }

    // ---- xtc frames (include/groan_xtc.h; XtcReader xtc_io/mod.rs:97-330, GroupXtcReader molly_xtc.rs:404-470)
    // frames [first, first + count) of an xtc file held in memory: the file's bytes are uploaded and decoded on the GPU
    void set_frames_xtc(const uint8_t *data, size_t len, const std::vector<uint64_t> &offsets, size_t first, size_t count) {
        if (first + count + 1 > offsets.size() || count == 0 || count > max_frames_) throw GpuError("Gpu", "frame range", GROAN_EINVAL);
        check(groan_gpu_push_xtc(ctx_, data, len, offsets.data() + first, count, nullptr, nullptr, nullptr), "set_frames_xtc");
        n_frames_ = count;
        touched();
    }
    // partial frames: xyz_sel holds only the atoms `atoms` (ascending) of every frame; the others keep their previous values
    void set_group_frames(const float *xyz_sel, const std::vector<uint32_t> &atoms, const SimBox *boxes, size_t n_frames) {
        std::vector<float> b(n_frames * 9);
        for (size_t f = 0; f < n_frames; f++) std::copy(boxes[f].m, boxes[f].m + 9, b.begin() + f * 9);
        check(groan_gpu_push_group_frames(ctx_, xyz_sel, atoms.data(), atoms.size(), b.data(), n_frames), "set_group_frames");
        n_frames_ = n_frames;
        touched();
    }
    // the current batch as xtc frames, byte for byte what XtcWriter::write_frame would append (xtc_io/mod.rs:300-330)
    std::vector<uint8_t> write_xtc(float precision, const std::vector<int32_t> &step, const std::vector<float> &time, int n_threads = 4) {
        std::vector<uint8_t> out(n_frames_ * (n_atoms_ * 12 + 128) + 64);
        size_t len = 0;
        check(groan_gpu_write_xtc(ctx_, precision, step.empty() ? nullptr : step.data(), time.empty() ? nullptr : time.data(), n_threads,
                                  out.data(), out.size(), &len), "write_xtc");
        out.resize(len);
        return out;
    }

    // ---- System::guess_bonds (guess.rs:362-395): bonds i < j with distance < (vdw[i] + vdw[j]) * radius_factor, per frame, sorted.
    // vdw[i] < 0: the atom has no van der Waals radius.  The bonds of frame 0 become the topology (assign_bonds, guess.rs:409-425).
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> guess_bonds(const std::vector<float> &vdw, float radius_factor = 0.55f,
                                                                        bool assign = true) {
        if (vdw.size() != n_atoms_) throw GpuError("Gpu", "one radius per atom", GROAN_EINVAL);
        std::vector<uint64_t> count(n_frames_, 0);
        size_t cap = 8 * n_atoms_ + 64;
        std::vector<uint32_t> pairs;
        for (int attempt = 0; attempt < 2; attempt++) {
            pairs.assign(n_frames_ * cap * 2, 0);
            check(groan_gpu_guess_bonds(ctx_, vdw.data(), radius_factor, count.data(), pairs.data(), cap), "guess_bonds");
            const uint64_t most = *std::max_element(count.begin(), count.end());
            if (most <= cap) break;
            cap = (size_t)most;
        }
        std::vector<std::vector<std::pair<uint32_t, uint32_t>>> out(n_frames_);
        for (size_t f = 0; f < n_frames_; f++) {
            for (uint64_t k = 0; k < count[f]; k++) out[f].emplace_back(pairs[(f * cap + k) * 2], pairs[(f * cap + k) * 2 + 1]);
            std::sort(out[f].begin(), out[f].end());
        }
        if (assign && n_frames_) add_bonds(out[0]);
        return out;
    }

    // ---- HBondAnalysis::analyze_single (hbonds.rs:240-320) for one acceptor group and one list of donors with their hydrogens
    struct HBond {
        uint32_t donor, hydrogen, acceptor;
        float distance, angle;
    };
    struct Donor {
        uint32_t donor;
        std::vector<uint32_t> hydrogens;
    };
    std::vector<std::vector<HBond>> hbonds_single(const std::string &acceptors, const std::vector<Donor> &donors, float max_distance,
                                                  float min_angle) {
        std::vector<uint32_t> don, off(1, 0), hyd;
        for (const Donor &d : donors) {
            don.push_back(d.donor);
            hyd.insert(hyd.end(), d.hydrogens.begin(), d.hydrogens.end());
            off.push_back((uint32_t)hyd.size());
        }
        if (hyd.empty()) hyd.push_back(0);
        std::vector<uint64_t> count(n_frames_, 0);
        size_t cap = 4 * hyd.size() + 64;
        std::vector<uint32_t> dha;
        std::vector<float> da;
        for (int attempt = 0; attempt < 2; attempt++) {
            dha.assign(n_frames_ * cap * 3, 0);
            da.assign(n_frames_ * cap * 2, 0.0f);
            check(groan_gpu_hbonds(ctx_, group(acceptors).gid, don.data(), off.data(), hyd.data(), don.size(), max_distance, min_angle,
                                   count.data(), dha.data(), da.data(), cap), "hbonds", acceptors);
            const uint64_t most = *std::max_element(count.begin(), count.end());
            if (most <= cap) break;
            cap = (size_t)most;
        }
        std::vector<std::vector<HBond>> out(n_frames_);
        for (size_t f = 0; f < n_frames_; f++) {
            for (uint64_t k = 0; k < count[f]; k++) {
                const size_t at = f * cap + k;
                out[f].push_back({dha[at * 3], dha[at * 3 + 1], dha[at * 3 + 2], da[at * 2], da[at * 2 + 1]});
            }
            std::sort(out[f].begin(), out[f].end(), [](const HBond &a, const HBond &b) {
                return std::tie(a.donor, a.acceptor, a.hydrogen) < std::tie(b.donor, b.acceptor, b.hydrogen);
            });
        }
        return out;
    }
    const std::vector<std::pair<uint32_t, uint32_t>> &bonds() const { return bonds_; }
    std::vector<uint32_t> group_indices(const std::string &name) const { return group(name).indices; }

    // diagnostics of the last centre / RMSD call (groan_gpu.h)
    size_t fallback_frames() {
        size_t n = 0;
        check(groan_gpu_fallback_frames(ctx_, &n), "fallback_frames");
        return n;
    }
    size_t second_pass_frames() {
        size_t n = 0;
        check(groan_gpu_second_pass_frames(ctx_, &n), "second_pass_frames");
        return n;
    }

    groan_gpu_ctx *raw() { return ctx_; }

  private:
    struct Group {
        int gid = -2;
        std::vector<uint32_t> indices;
    };

    const Group &group(const std::string &name, bool rmsd = false) const {
        auto it = groups_.find(name);
        if (it == groups_.end()) {
            if (rmsd) throw RMSDError("NonexistentGroup", name, GROAN_ENOGROUP);
            throw GroupError("NotFound", name, GROAN_ENOGROUP);
        }
        return it->second;
    }
    void upload_group(const Group &g) {
        std::vector<float> m;
        if (!masses_.empty()) {
            m.reserve(g.indices.size());
            for (uint32_t i : g.indices) m.push_back(masses_[i]);
        }
        check(groan_gpu_set_group(ctx_, g.gid, g.indices.data(), g.indices.size(), masses_.empty() ? nullptr : m.data()), "group_create");
    }
    std::vector<Vector3D> centre(const std::string &name, int kind, int weighted) {
        std::vector<Vector3D> out(n_frames_);
        const int gid = group(name).gid;
        int st;
        if (kind == 0) st = groan_gpu_estimate_center(ctx_, gid, weighted, &out[0][0]);
        else if (kind == 1) st = groan_gpu_get_center(ctx_, gid, weighted, &out[0][0]);
        else st = groan_gpu_get_center_naive(ctx_, gid, &out[0][0]);
        check(st, "group_center", name);
        return out;
    }
    void set_reference(const System &ref, const std::string &name) {
        const Group &mine = group(name, true);
        const Group &theirs = ref.group(name, true);
        const auto key = std::make_pair(&ref, ref.version_);
        auto it = ref_keys_.find(name);
        if (it != ref_keys_.end() && it->second == key) return;
        if (ref.frame0_.empty()) throw GpuError("NoFrames", "the reference system has no frame", GROAN_ENOFRAMES);
        if (ref.frame0_stale_) { // the reference was wrapped / translated / fitted in place since it was uploaded
            std::vector<float> all(ref.n_frames_ * ref.n_atoms_ * 3);
            ref.check(groan_gpu_get_frames(ref.ctx_, all.data()), "calc_rmsd (reading the reference back)");
            ref.frame0_.assign(all.begin(), all.begin() + ref.n_atoms_ * 3);
            ref.frame0_stale_ = false;
        }
        const bool all_atoms = theirs.gid == GROAN_GROUP_ALL;
        const size_t n_ref = all_atoms ? ref.n_atoms_ : theirs.indices.size();
        std::vector<float> m;
        if (!ref.masses_.empty()) {
            if (all_atoms) m = ref.masses_;
            else for (uint32_t i : theirs.indices) m.push_back(ref.masses_[i]);
        }
        check(groan_gpu_rmsd_set_reference(ctx_, mine.gid, ref.frame0_.data(), ref.n_atoms_, all_atoms ? nullptr : theirs.indices.data(),
                                           n_ref, ref.has_box0_ ? ref.box0_.m : nullptr, m.empty() ? nullptr : m.data()),
              "calc_rmsd", name, true);
        ref_keys_[name] = key;
    }

    // status -> the reference's error enums (same mapping as INTEGRATION.md section 3)
    void check(int st, const char *what, const std::string &grp = std::string(), bool rmsd = false) const {
        if (st == GROAN_OK) return;
        const std::string msg = std::string(groan_gpu_strerror(st)) + " in " + what + (grp.empty() ? "" : " (" + grp + ")");
        size_t a = 0, b = 0;
        groan_gpu_error_detail(ctx_, &a, &b);
        switch (st) {
        case GROAN_ENOBOX:
        case GROAN_ENOTORTHO:
        case GROAN_EZEROBOX: {
            const std::string v = st == GROAN_ENOBOX ? "SimBoxError::DoesNotExist"
                                  : st == GROAN_ENOTORTHO ? "SimBoxError::NotOrthogonal" : "panic: Box len should not be zero";
            if (rmsd) throw RMSDError("InvalidSimBox(" + v + ")", msg, st);
            if (!grp.empty()) throw GroupError("InvalidSimBox(" + v + ")", msg, st);
            throw SimBoxError(v, msg, st);
        }
        case GROAN_EEMPTY:
            if (rmsd) throw RMSDError("EmptyGroup", msg, st);
            throw GroupError("EmptyGroup", msg, st);
        case GROAN_ENOGROUP:
            if (rmsd) throw RMSDError("NonexistentGroup", msg, st);
            throw GroupError("NotFound", msg, st);
        case GROAN_ENOPOS: {
            const std::string v = "InvalidPosition(PositionError::NoPosition(" + std::to_string(b) + "))";
            if (rmsd) {
                RMSDError r(v, msg, st);
                r.a = a;
                r.b = b;
                throw r;
            }
            GroupError ge(v, msg, st);
            ge.a = a;
            ge.b = b;
            throw ge;
        }
        case GROAN_ENOMASS: {
            const std::string v = "InvalidMass(MassError::NoMass(" + std::to_string(b) + "))";
            if (rmsd) throw RMSDError(v, msg, st);
            throw GroupError(v, msg, st);
        }
        case GROAN_EGROUPSIZE: {
            RMSDError e("InconsistentGroup", msg + ": " + std::to_string(a) + " atoms in the reference, " + std::to_string(b) + " in the target", st);
            e.a = a;
            e.b = b;
            throw e;
        }
        case GROAN_ECUDA: throw GpuError("CudaError", groan_gpu_last_cuda_error(ctx_), st);
        default:
            throw GpuError(st == GROAN_EINVAL ? "InvalidArgument" : st == GROAN_ENOFRAMES ? "NoFrames" : st == GROAN_ENOREF ? "NoReference"
                           : st == GROAN_ECAPACITY ? "Capacity" : "Unknown", msg, st);
        }
    }

    groan_gpu_ctx *ctx_ = nullptr;
    size_t n_atoms_, max_frames_, n_frames_ = 0;
    std::vector<std::pair<uint32_t, uint32_t>> bonds_;  // topology given through add_bonds / guess_bonds (i < j, sorted)
    int next_gid_ = 0;
    std::map<std::string, Group> groups_;
    std::vector<float> masses_;
    std::vector<uint32_t> mol_ref_;
    // positions changed on the device: the host copy of frame 0 is stale and RMSD references taken from this System are too
    void touched() {
        frame0_stale_ = true;
        version_++;
    }
    mutable std::vector<float> frame0_;
    mutable bool frame0_stale_ = false;
    SimBox box0_;
    bool has_box0_ = false;
    unsigned long version_ = 0;
    std::map<std::string, std::pair<const System *, unsigned long>> ref_keys_;
};

// src/system/hbonds.rs: chains of (acceptors, donors, hydrogens) given as groups of the System (the reference takes selection
// queries; the selection language stays outside this library), pairs of chains to analyse, the two criteria.
struct HBondChain {
    std::string acceptors, donors, hydrogens;
};
class HBondAnalysis {
  public:
    using HBondMap = std::map<std::pair<size_t, size_t>, std::vector<System::HBond>>;
    // HBondAnalysis construction = HBondChainGroups::new per chain (hbonds.rs:108-150: hydrogens of a donor = its bonded atoms
    // that are in the hydrogen group; donors without one are dropped) + sanity_check_pairs (:337-370)
    HBondAnalysis(System &system, const std::vector<HBondChain> &chains, std::vector<std::pair<size_t, size_t>> pairs, float max_distance,
                  float min_angle)
        : sys_(system), pairs_(std::move(pairs)), max_distance_(max_distance), min_angle_(min_angle) {
        std::map<uint32_t, std::vector<uint32_t>> bonded;
        for (const auto &b : system.bonds()) {
            bonded[b.first].push_back(b.second);
            bonded[b.second].push_back(b.first);
        }
        for (const HBondChain &c : chains) {
            const std::vector<uint32_t> hyd = system.group_indices(c.hydrogens);
            Chain ch;
            ch.acceptors = c.acceptors;
            for (uint32_t d : system.group_indices(c.donors)) {
                System::Donor don{d, {}};
                std::vector<uint32_t> nb = bonded[d];
                std::sort(nb.begin(), nb.end());
                for (uint32_t h : nb)
                    if (std::binary_search(hyd.begin(), hyd.end(), h)) don.hydrogens.push_back(h);
                if (!don.hydrogens.empty()) ch.donors.push_back(std::move(don));
            }
            if (system.group_get_n_atoms(c.acceptors) == 0 && ch.donors.empty()) throw GroanError("EmptyChain", "hydrogen-bond chain is empty", GROAN_EINVAL);
            chains_.push_back(std::move(ch));
        }
        std::vector<std::pair<size_t, size_t>> seen;
        for (const auto &p : pairs_) {
            if (p.first >= chains_.size() || p.second >= chains_.size()) throw GroanError("InvalidPair", "pair names a chain that does not exist", GROAN_EINVAL);
            const std::pair<size_t, size_t> key(std::min(p.first, p.second), std::max(p.first, p.second));
            if (std::find(seen.begin(), seen.end(), key) != seen.end()) throw GroanError("DuplicatePair", "pair requested twice", GROAN_EINVAL);
            seen.push_back(key);
        }
    }
    // FrameAnalyze::analyze (hbonds.rs:160-210) for every frame of the System's current batch
    std::vector<HBondMap> analyze() {
        std::vector<HBondMap> maps(sys_.n_frames());
        for (const auto &p : pairs_) {
            std::vector<std::vector<std::vector<System::HBond>>> parts;
            if (p.first == p.second) {
                parts.push_back(sys_.hbonds_single(chains_[p.first].acceptors, chains_[p.first].donors, max_distance_, min_angle_));
            } else {  // analyze_pair (hbonds.rs:214-238)
                parts.push_back(sys_.hbonds_single(chains_[p.first].acceptors, chains_[p.second].donors, max_distance_, min_angle_));
                parts.push_back(sys_.hbonds_single(chains_[p.second].acceptors, chains_[p.first].donors, max_distance_, min_angle_));
            }
            for (size_t f = 0; f < sys_.n_frames(); f++)
                for (auto &part : parts) {
                    auto &dst = maps[f][p];
                    dst.insert(dst.end(), part[f].begin(), part[f].end());
                }
        }
        return maps;
    }

  private:
    struct Chain {
        std::string acceptors;
        std::vector<System::Donor> donors;
    };
    System &sys_;
    std::vector<Chain> chains_;
    std::vector<std::pair<size_t, size_t>> pairs_;
    float max_distance_, min_angle_;
};

// An xtc trajectory held in memory (std::ifstream / mmap of the file): frame offsets from groan_xtc_scan
struct XtcFile {
    std::vector<uint8_t> data;
    std::vector<uint64_t> offsets;  // n_frames + 1
    int32_t n_atoms = 0;
    explicit XtcFile(std::vector<uint8_t> bytes) : data(std::move(bytes)) {
        size_t cap = 1024, n = 0;
        for (;;) {
            offsets.assign(cap + 1, 0);
            const int st = groan_xtc_scan(data.data(), data.size(), cap, offsets.data(), &n_atoms, &n);
            if (st != GROAN_XTC_OK && st != GROAN_XTC_EOF) throw GpuError("Xtc", "not an xtc file", st);
            if (n < cap) break;
            cap *= 8;
        }
        offsets.resize(n + 1);
    }
    size_t n_frames() const { return offsets.size() - 1; }
};


// The seam the reference offers to per-frame code is `body: Fn(&System, &mut Data)` of traj_iter_map_reduce
// (parallel.rs:208-269) / FrameAnalyze::analyze (traj_convert.rs:76-83): one frame at a time.  A GPU operator wants batches:
// FrameBatcher collects the frames handed to it one by one and, every `batch` frames (and at finish(), the reducer's
// place), stages them with one set_frames and calls `flush(system, first_frame_index, n)`.
class FrameBatcher {
  public:
    using Flush = std::function<void(System &, size_t first_frame, size_t n_frames)>;
    FrameBatcher(System &system, size_t batch, Flush flush) : sys_(system), batch_(batch), flush_(std::move(flush)) {
        xyz_.reserve(batch * system.get_n_atoms() * 3);
        boxes_.reserve(batch);
    }
    // what the body closure does with the frame the reader has just produced
    void push(const float *xyz, const SimBox &box) {
        xyz_.insert(xyz_.end(), xyz, xyz + sys_.get_n_atoms() * 3);
        boxes_.push_back(box);
        if (boxes_.size() == batch_) flush();
    }
    void finish() {
        if (!boxes_.empty()) flush();
    }
    size_t frames_seen() const { return seen_; }

  private:
    void flush() {
        sys_.set_frames(xyz_.data(), boxes_.data(), boxes_.size());
        flush_(sys_, seen_, boxes_.size());
        seen_ += boxes_.size();
        xyz_.clear();
        boxes_.clear();
    }
    System &sys_;
    size_t batch_, seen_ = 0;
    Flush flush_;
    std::vector<float> xyz_;
    std::vector<SimBox> boxes_;
};

} // namespace groan
