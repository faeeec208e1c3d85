// loaders.cpp — the loaders' hot spot that stands between a 100 k-atom file and its first frame.
//
// PDBReader::loadAtomsFromFile (solr/io/PDBReader.cpp:616-660) finds the bonds ("sticks") of a molecule by testing every atom against
// every other atom: for each atom, in std::map order of the atom ids, every OTHER atom with processed < 2 and the same backbone flag
// whose distance is below the stick distance (DEFAULT_STICK_DISTANCE = 1.7, doubled between backbone atoms for the backbone
// geometry) gets a cylinder, in map order again.  That is n^2 distance tests — 10^10 for the 100 k atoms of config 2 — before a
// single primitive exists.  b200h_find_bonds returns the same pairs in the same order from a uniform grid whose cell edge is the
// largest stick distance: an atom's partners can only sit in the 27 cells around it; the candidates are tested with the reference's
// own float expression (sqrtf(dx*dx + dy*dy + dz*dz) < stickDistance, one rounding per operation), so the grid changes which pairs
// are LOOKED AT, never which are taken.  The loader keeps its loop over the atoms and replaces the inner loop by a walk over
// pairs[first[i] .. first[i + 1]) (INTEGRATION.md has the three lines).
//
// OBJReader keeps its vertices in std::map<int, vec3f> keyed 1..n (OBJReader.cpp:415-418): a vector indexed by the key is the same
// container for that use and needs no code of its own.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {
// atoms in std::map iteration order (ascending id): xyz[3 n] positions as the file gives them, processed[n], isBackbone[n].
// backboneGeometry: geometryType == gtBackbone.  Writes, per atom i, the indices j of its partners in ascending order:
// partners[first[i] .. first[i + 1]); first has n + 1 entries.  Returns the number of pairs, or -(needed capacity) if
// partnerCapacity is too small (nothing but `first` is written then).
long b200h_find_bonds(const float* xyz, const int* processed, const unsigned char* isBackbone, int n, int backboneGeometry,
                      float stickDistance, int* first, int* partners, long partnerCapacity)
{
    if (n <= 0) { if (first) first[0] = 0; return 0; }
    const float maxStick = backboneGeometry ? stickDistance * 2.f : stickDistance;
    float lo[3] = {xyz[0], xyz[1], xyz[2]}, hi[3] = {xyz[0], xyz[1], xyz[2]};
    for (int i = 1; i < n; ++i)
        for (int a = 0; a < 3; ++a)
        {
            lo[a] = std::min(lo[a], xyz[3 * (size_t)i + a]);
            hi[a] = std::max(hi[a], xyz[3 * (size_t)i + a]);
        }
    // cells a little larger than the largest stick distance: a partner is never more than one cell away
    const double cell = (double)maxStick * 1.0001 + 1e-6;
    int dim[3];
    for (int a = 0; a < 3; ++a)
    {
        const double cells = std::floor(((double)hi[a] - lo[a]) / cell) + 1.0;
        dim[a] = (int)std::min(cells, 1024.0);
    }
    auto cellOf = [&](const float* p, int c[3]) {
        for (int a = 0; a < 3; ++a)
        {
            int k = (int)std::floor(((double)p[a] - lo[a]) / cell);
            c[a] = k < 0 ? 0 : (k >= dim[a] ? dim[a] - 1 : k); // a clamped axis (more than 1024 cells) only merges cells: still conservative
        }
    };
    const size_t nbCells = (size_t)dim[0] * dim[1] * dim[2];
    std::vector<int> start(nbCells + 1, 0), order(n);
    std::vector<int> cellIndex(n);
    for (int i = 0; i < n; ++i)
    {
        int c[3];
        cellOf(xyz + 3 * (size_t)i, c);
        cellIndex[i] = (c[0] * dim[1] + c[1]) * dim[2] + c[2];
        ++start[cellIndex[i] + 1];
    }
    for (size_t k = 0; k < nbCells; ++k) start[k + 1] += start[k];
    {
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int i = 0; i < n; ++i) order[fill[cellIndex[i]]++] = i; // ascending i inside a cell
    }
    // when an axis was clamped its last cell is wider than `cell`: partners of an atom there may be farther than one cell in the
    // unclamped grid but are still in the same (merged) or the neighbouring cell of the clamped one
    long total = 0;
    std::vector<int> mine;
    for (int pass = 0; pass < 2; ++pass)
    {
        long at = 0;
        for (int i = 0; i < n; ++i)
        {
            if (pass == 0) first[i] = (int)at;
            mine.clear();
            int c[3];
            cellOf(xyz + 3 * (size_t)i, c);
            const float* pi = xyz + 3 * (size_t)i;
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dz = -1; dz <= 1; ++dz)
                    {
                        const int x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
                        if (x < 0 || y < 0 || z < 0 || x >= dim[0] || y >= dim[1] || z >= dim[2]) continue;
                        const size_t k = ((size_t)x * dim[1] + y) * dim[2] + z;
                        for (int s = start[k]; s < start[k + 1]; ++s)
                        {
                            const int j = order[s];
                            // PDBReader.cpp:621-634, the reference's own test
                            if (j == i || !(processed[j] < 2) || isBackbone[i] != isBackbone[j]) continue;
                            const float* pj = xyz + 3 * (size_t)j;
                            const float ax = pi[0] - pj[0], ay = pi[1] - pj[1], az = pi[2] - pj[2];
                            const float distance = sqrtf(ax * ax + ay * ay + az * az);
                            const float limit = (backboneGeometry && isBackbone[j]) ? stickDistance * 2.f : stickDistance;
                            if (distance < limit) mine.push_back(j);
                        }
                    }
            if (pass == 0) { at += (long)mine.size(); continue; }
            std::sort(mine.begin(), mine.end());
            memcpy(partners + at, mine.data(), mine.size() * sizeof(int));
            at += (long)mine.size();
        }
        if (pass == 0)
        {
            total = at;
            first[n] = (int)total;
            if (!partners || total > partnerCapacity) return total > partnerCapacity ? -total : total;
        }
    }
    return total;
}
} // extern "C"
