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
// OBJReader::loadModelFromFile reads the file twice (solr/io/OBJReader.cpp:440-563 the "v" / "vn" / "vt" lines, :602 onwards the
// faces).  Its first pass keeps the vertices, normals and texture coordinates in three std::map<int, ...> keyed 1..n
// (OBJReader.cpp:415-418) and builds every number character by character in a std::string: for the 1 M-triangle mesh of config 3
// (500 k vertices + normals) that is 1.5 M map nodes and 30 M string appends before a single face is read.  b200h_obj_vertex_pass is
// that first pass over the whole text at once into flat arrays indexed by key - 1 — the same tokenisation (a separator is a blank
// that follows a non-blank; later blanks join the next token, which atof then skips), the same atof, the same sign flips
// (z of vertices and normals negated), the same folding of negative texture coordinates, the same bounding box — so the second
// pass reads `vertices[3 * (key - 1)]` where it read `vertices[key]`.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
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
// First pass of OBJReader::loadModelFromFile over the text of a .obj file (OBJReader.cpp:440-563).  vertices / normals: 3 floats
// per entry, texCoords: 2; any of them may be NULL (counting call).  counts[3] = entries found (vertices, normals, texture
// coordinates); aabb[6] = min xyz, max xyz of the vertices, starting from the reference's +/-100000.  Entries beyond a capacity are
// counted but not stored.  Returns 0.
int b200h_obj_vertex_pass(const char* text, size_t len, float* vertices, int vertexCapacity, float* normals, int normalCapacity,
                          float* texCoords, int texCoordCapacity, int* counts, float* aabb)
{
    int nv = 0, nn = 0, nt = 0;
    float box[6] = {100000.f, 100000.f, 100000.f, -100000.f, -100000.f, -100000.f};
    std::vector<char> line;  // the line without its carriage returns (:443)
    char token[512];
    size_t at = 0;
    while (at < len)
    {
        size_t end = at;
        while (end < len && text[end] != '\n') ++end;
        line.clear();
        for (size_t k = at; k < end; ++k)
            if (text[k] != '\r') line.push_back(text[k]);
        at = end + 1;
        const size_t n = line.size();
        if (n <= 1 || line[0] != 'v') continue;
        float v[3] = {0.f, 0.f, 0.f};
        // :457-512 — a blank after a non-blank closes an item; item 0 is the keyword itself; items 1..3 are x, y, z
        size_t i = 1, tl = 0;
        int item = 0;
        char previous = line[0];
        auto assign = [&](int it) {
            if (it >= 1 && it <= 3) { token[tl] = 0; v[it - 1] = static_cast<float>(atof(token)); }
        };
        while (i < n && item < 4)
        {
            if (line[i] == ' ' && previous != ' ')
            {
                assign(item);
                ++item;
                tl = 0;
            }
            else if (tl + 1 < sizeof(token)) token[tl++] = line[i];
            previous = line[i];
            ++i;
        }
        if (tl != 0) assign(item);
        if (line[1] == 'n')
        {
            if (normals && nn < normalCapacity) { normals[3 * (size_t)nn] = v[0]; normals[3 * (size_t)nn + 1] = v[1]; normals[3 * (size_t)nn + 2] = -v[2]; }
            ++nn;
        }
        else if (line[1] == 't')
        {
            float x = v[0], y = v[1];
            if (x < 0.f) { const int a = static_cast<int>(std::fabs(x)); x = std::fabs(x) - a; }
            if (y < 0.f) { const int a = static_cast<int>(std::fabs(y)); y = std::fabs(y) - a; }
            if (texCoords && nt < texCoordCapacity) { texCoords[2 * (size_t)nt] = x; texCoords[2 * (size_t)nt + 1] = y; }
            ++nt;
        }
        else if (line[1] == ' ')
        {
            const float z = -v[2];
            if (vertices && nv < vertexCapacity) { vertices[3 * (size_t)nv] = v[0]; vertices[3 * (size_t)nv + 1] = v[1]; vertices[3 * (size_t)nv + 2] = z; }
            ++nv;
            box[0] = (v[0] < box[0]) ? v[0] : box[0]; box[1] = (v[1] < box[1]) ? v[1] : box[1]; box[2] = (z < box[2]) ? z : box[2];
            box[3] = (v[0] > box[3]) ? v[0] : box[3]; box[4] = (v[1] > box[4]) ? v[1] : box[4]; box[5] = (z > box[5]) ? z : box[5];
        }
    }
    if (counts) { counts[0] = nv; counts[1] = nn; counts[2] = nt; }
    if (aabb) memcpy(aabb, box, sizeof(box));
    return 0;
}
} // extern "C"
