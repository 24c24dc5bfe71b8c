// Parity tests through the C++ mirror (include/groan_gpu.hpp), written like the reference's own unit tests.
// Each block cites the reference test it restates.  Exit code 0 = all passed.  Needs a GPU (pytest -m gpu runs it).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "groan_gpu.hpp"

using namespace groan;

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                    \
        }                                                                  \
    } while (0)
static bool approx(float a, float b, float eps = 1e-5f) { return std::fabs(a - b) <= eps; }
static bool approx3(const Vector3D &v, float x, float y, float z, float eps = 1e-5f) {
    return approx(v[0], x, eps) && approx(v[1], y, eps) && approx(v[2], z, eps);
}

int main(int argc, char **argv) {
    const std::string golden = argc > 1 ? argv[1] : "tests/golden";
    // ---- analysis.rs:488-560 (center_single_atom / center_two_atoms / center_several_atoms), box 10^3
    {
        System s(5);
        s.set_masses({10.3f, 5.4f, 3.8f, 10.1f, 7.6f});
        s.group_create_from_indices("g", {0, 1, 2, 3, 4});
        const std::vector<Vector3D> five = {{3.3f, 0.3f, 2.5f}, {4.3f, 1.2f, 9.8f}, {3.2f, 5.6f, 0.5f}, {0.2f, 9.0f, 6.6f}, {8.7f, 5.0f, 2.4f}};
        s.set_frame(five, SimBox::orthogonal(10, 10, 10));
        CHECK(approx3(s.group_estimate_center("g")[0], 2.634386f, 9.775156f, 1.1748f, 1e-4f));
        CHECK(approx3(s.group_estimate_com("g")[0], 1.9526f, 9.7567f, 1.8812f, 1e-4f)); // analysis.rs:930-988
        // the same atoms moved by whole box vectors (analysis.rs:562-629)
        const std::vector<Vector3D> out = {{3.3f, 10.3f, 2.5f}, {4.3f, 1.2f, -0.2f}, {13.2f, 15.6f, 0.5f}, {10.2f, -1.0f, 6.6f}, {-1.3f, 5.0f, 2.4f}};
        s.set_frame(out, SimBox::orthogonal(10, 10, 10));
        CHECK(approx3(s.group_estimate_center("g")[0], 2.634386f, 9.775156f, 1.1748f, 1e-4f));
    }
    {
        System p(2);
        p.set_masses({12.8f, 0.4f});
        p.group_create_from_indices("g", {0, 1});
        p.set_frame({{4.5f, 3.2f, 1.7f}, {9.8f, 9.5f, 3.0f}}, SimBox::orthogonal(10, 10, 10));
        CHECK(approx3(p.group_get_center("g")[0], 2.15f, 1.35f, 2.35f));
        CHECK(approx3(p.group_get_com("g")[0], 4.3575745f, 3.0878792f, 1.7393947f));
        // vector3d.rs:1040-1207 through Atom::distance: box 4^3 in the reference; here the two atoms in box 10^3
        const float dx = p.group_all_distances("g", "g", Dimension::X)[1]; // D[0,1] = signed x distance of atom 0 from atom 1
        CHECK(approx(dx, 4.5f - 9.8f + 10.0f));
        const float dxyz = p.group_all_distances("g", "g", Dimension::XYZ)[1];
        CHECK(approx(dxyz, std::sqrt(4.7f * 4.7f + 3.7f * 3.7f + 1.3f * 1.3f)));
    }
    // ---- vector3d.rs:1017-1037 (wrap): box 2^3
    {
        System w(3);
        w.set_frame({{-1.0f, 1.5f, 3.0f}, {2.0f, 2.2f, -0.3f}, {-54.2f, 77.8f, 124.5f}}, SimBox::orthogonal(2, 2, 2));
        w.atoms_wrap();
        const std::vector<float> x = w.get_frames();
        CHECK(approx(x[0], 1.0f) && approx(x[1], 1.5f) && approx(x[2], 1.0f));
        CHECK(x[3] == 2.0f && approx(x[4], 0.2f) && approx(x[5], 1.7f)); // x == L stays L: the strict comparison of wrap_coordinate
        CHECK(approx(x[6], 1.8f, 1e-4f) && approx(x[7], 1.8f, 1e-4f) && approx(x[8], 0.5f, 1e-4f));
    }
    // ---- rmsd.rs: identical and rigidly translated structures have RMSD 0; a known displacement has a known RMSD
    {
        const size_t n = 64;
        std::vector<Vector3D> ref(n), moved(n), bumped(n);
        for (size_t i = 0; i < n; i++) {
            ref[i] = {2.0f + 0.1f * float(i % 8), 3.0f + 0.1f * float((i / 8) % 8), 4.0f + 0.05f * float(i % 5)};
            moved[i] = {ref[i][0] + 7.25f, ref[i][1] - 3.5f, ref[i][2] + 9.0f}; // across the periodic boundary
            bumped[i] = ref[i];
        }
        bumped[0][2] += 0.8f; // one atom of 64 displaced: small, known RMSD after the fit
        System a(n, 3), r(n);
        std::vector<float> m(n, 12.0f);
        a.set_masses(m);
        r.set_masses(m);
        std::vector<uint32_t> idx(n);
        for (size_t i = 0; i < n; i++) idx[i] = (uint32_t)i;
        a.group_create_from_indices("G", idx);
        r.group_create_from_indices("G", idx);
        const SimBox box = SimBox::orthogonal(10, 10, 10);
        r.set_frame(ref, box);
        std::vector<float> xyz;
        for (const auto *fr : {&ref, &moved, &bumped})
            for (const auto &v : *fr) xyz.insert(xyz.end(), v.begin(), v.end());
        const SimBox boxes[3] = {box, box, box};
        a.set_frames(xyz.data(), boxes, 3);
        const std::vector<float> rm = a.calc_rmsd(r, "G");
        CHECK(approx(rm[0], 0.0f, 1e-4f) && approx(rm[1], 0.0f, 1e-4f));
        CHECK(rm[2] > 0.05f && rm[2] < 0.8f / std::sqrt(float(n)) + 1e-3f); // the fit can only reduce sqrt(0.8^2 / 64) = 0.1
        const auto cr = a.group_center_and_rmsd(r, "G");
        CHECK(approx(cr.second[2], rm[2], 2e-6f));
        // RMSDError::InconsistentGroup (rmsd.rs:1110-1130), NonexistentGroup
        r.group_create_from_indices("H", {0, 1, 2});
        a.group_create_from_indices("H", {0, 1});
        try {
            a.calc_rmsd(r, "H");
            CHECK(false);
        } catch (const RMSDError &e) {
            CHECK(e.variant == "InconsistentGroup" && e.a == 3 && e.b == 2);
        }
        try {
            a.calc_rmsd(r, "Nope");
            CHECK(false);
        } catch (const RMSDError &e) {
            CHECK(e.variant == "NonexistentGroup");
        }
    }
    // ---- modifying.rs:1079-1106 (make_group_whole on an artificial system) and the error enums
    {
        System s(3);
        s.set_frame({{1.0f, 4.0f, 2.0f}, {4.0f, 1.0f, 2.0f}, {1.0f, 1.0f, 2.0f}}, SimBox::orthogonal(5, 5, 5));
        s.make_group_whole("all");
        const std::vector<float> x = s.get_frames();
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) CHECK(std::fabs(x[i * 3 + k] - x[j * 3 + k]) <= 2.5f + 1e-5f);
        s.add_bonds({{0, 1}, {1, 2}});
        s.atoms_translate({3.5f, 4.5f, -3.0f});
        s.make_molecules_whole();
        const std::vector<float> y = s.get_frames();
        CHECK(y[0] >= 0 && y[0] <= 5 && y[1] >= 0 && y[1] <= 5 && y[2] >= 0 && y[2] <= 5); // the reference atom is in the box
        for (int i = 1; i < 3; i++)
            for (int k = 0; k < 3; k++) CHECK(std::fabs(y[i * 3 + k] - y[k]) <= 2.5f + 1e-5f);
        try {
            s.group_get_center("Nonexistent");
            CHECK(false);
        } catch (const GroupError &e) {
            CHECK(e.variant == "NotFound");
        }
        s.group_create_from_indices("Empty", {});
        try {
            s.group_get_center("Empty");
            CHECK(false);
        } catch (const GroupError &e) {
            CHECK(e.variant == "EmptyGroup");
        }
        const std::vector<float> again = s.get_frames();
        s.set_frames(again.data(), nullptr, 1);
        try {
            s.atoms_center("all", Dimension::XY);
            CHECK(false);
        } catch (const GroupError &e) {
            CHECK(e.variant.find("DoesNotExist") != std::string::npos);
        }
    }
    // ---- cellgrid.rs: neighbours within a cutoff == the pairs of the all-pairs matrix below it
    {
        const size_t n = 3000;
        System s(n);
        std::vector<Vector3D> x(n);
        unsigned long long st = 12345;
        auto rnd = [&]() {
            st = st * 6364136223846793005ULL + 1442695040888963407ULL;
            return float((st >> 33) & 0xFFFFFF) / float(0x1000000);
        };
        for (auto &v : x) v = {rnd() * 6.0f, rnd() * 5.0f, rnd() * 7.0f};
        s.set_frame(x, SimBox::orthogonal(6, 5, 7));
        s.group_create_from_range("A", 0, 99);
        s.group_create_from_range("B", 50, 2999);
        const float cutoff = 0.8f;
        const std::vector<float> D = s.group_all_distances("A", "B", Dimension::XYZ);
        size_t expect = 0;
        for (float d : D) expect += d < cutoff;
        const auto pl = s.group_pairs_within("A", "B", cutoff, expect + 8);
        CHECK(pl.count[0] == expect);
        bool ok = true;
        for (size_t k = 0; k < expect; k++) ok = ok && D[pl.pairs[k][0] * 2950 + pl.pairs[k][1]] == pl.dist[k] && pl.dist[k] < cutoff;
        CHECK(ok);
    }
    // ---- hbonds.rs:505-587 (test_hbonds_analyze_simple_water): water-water hydrogen bonds of aa_membrane_peptide.xtc, 0.3 nm / 150
    // degrees; the trajectory is read by the library's own xtc reader (file bytes uploaded, decoded on the GPU).  Water starts at
    // atom 17515 (OW, HW1, HW2 per molecule; aa_membrane_peptide.gro), 5 091 molecules.
    {
        std::ifstream in(golden + "/xtc/aa_membrane_peptide.xtc", std::ios::binary);
        std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        CHECK(!bytes.empty());
        if (!bytes.empty()) {
            XtcFile xf(std::move(bytes));
            CHECK(xf.n_atoms == 32817 && xf.n_frames() == 21);
            System s((size_t)xf.n_atoms, xf.n_frames());
            s.set_frames_xtc(xf.data.data(), xf.data.size(), xf.offsets, 0, xf.n_frames());
            std::vector<uint32_t> ow, hw;
            std::vector<std::pair<uint32_t, uint32_t>> bonds;
            for (uint32_t k = 0; k < 5091; k++) {
                const uint32_t o = 17515 + 3 * k;
                ow.push_back(o);
                hw.push_back(o + 1);
                hw.push_back(o + 2);
                bonds.emplace_back(o, o + 1);
                bonds.emplace_back(o, o + 2);
            }
            s.group_create_from_indices("OW", ow);
            s.group_create_from_indices("HW", hw);
            s.add_bonds(bonds);
            HBondAnalysis an(s, {{"OW", "OW", "HW"}}, {{0, 0}}, 0.3f, 150.0f);
            const auto maps = an.analyze();
            const size_t expected_n[21] = {4675, 4644, 4629, 4617, 4651, 4532, 4649, 4611, 4621, 4701, 4694, 4650, 4565, 4681,
                                           4699, 4711, 4652, 4649, 4697, 4652, 4644};
            for (size_t f = 0; f < 21; f++) CHECK(maps[f].at({0, 0}).size() == expected_n[f]);
            // the first bond of the reference's list for frame 0: HBond::new(17527, 17528, 21100, 0.262, 157.241)
            bool found = false;
            for (const auto &h : maps[0].at({0, 0}))
                if (h.donor == 17527 && h.hydrogen == 17528 && h.acceptor == 21100) found = approx(h.distance, 0.262f, 1e-3f) && approx(h.angle, 157.241f, 1e-3f);
            CHECK(found);
            // guess_bonds on the same frames: with the radii of elements.yaml (O 0.152, H 0.12 ... here only the water) every
            // water must come out as O-H, O-H and nothing else inside the water
            std::vector<float> vdw((size_t)xf.n_atoms, -1.0f);
            for (uint32_t o : ow) { vdw[o] = 0.152f; vdw[o + 1] = 0.12f; vdw[o + 2] = 0.12f; }
            const auto gb = s.guess_bonds(vdw, 0.55f, false);
            CHECK(gb.size() == 21 && gb[0].size() == 2 * 5091);
            bool all_oh = true;
            for (const auto &b : gb[0]) all_oh = all_oh && (b.first - 17515) % 3 == 0 && b.second > b.first && b.second <= b.first + 2;
            CHECK(all_oh);
            // write the batch back at the file's precision: decoding what we wrote gives the same lattice, so the bytes of the
            // coordinates section are the file's (headers carry step / time, passed through here as zeros)
            const std::vector<uint8_t> again = s.write_xtc(100.0f, {}, {});
            XtcFile back(again);
            CHECK(back.n_frames() == 21 && back.n_atoms == 32817);
            System s2((size_t)back.n_atoms, back.n_frames());
            s2.set_frames_xtc(back.data.data(), back.data.size(), back.offsets, 0, back.n_frames());
            CHECK(s.get_frames() == s2.get_frames());
        }
    }
    // ---- FrameBatcher: frames arrive one by one (traj_iter_map_reduce body), results come back per batch, in order
    {
        const size_t n = 5000, total = 11, batch = 4;
        System s(n, batch);
        s.group_create_from_range("Half", 0, 2499);
        std::vector<float> centres;
        FrameBatcher fb(s, batch, [&](System &sys, size_t first, size_t nf) {
            CHECK(first == centres.size());
            for (const Vector3D &c : sys.group_get_center("Half")) centres.push_back(c[0]);
            CHECK(nf == sys.n_frames());
        });
        std::vector<float> frame(n * 3);
        for (size_t f = 0; f < total; f++) {
            for (size_t i = 0; i < n; i++) {
                frame[i * 3 + 0] = 1.0f + 0.5f * float(f) + 0.001f * float(i % 100);
                frame[i * 3 + 1] = 3.0f;
                frame[i * 3 + 2] = 4.0f;
            }
            fb.push(frame.data(), SimBox::orthogonal(20, 20, 20));
        }
        fb.finish();
        CHECK(centres.size() == total && fb.frames_seen() == total);
        for (size_t f = 0; f < total; f++) CHECK(approx(centres[f], 1.0f + 0.5f * float(f) + 0.0495f, 2e-5f));
    }
    if (failures == 0) std::printf("cpp mirror: all checks passed\n");
    return failures == 0 ? 0 : 1;
}
