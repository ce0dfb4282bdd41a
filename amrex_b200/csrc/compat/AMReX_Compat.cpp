// Runtime half of the source compatibility layer (AMReX_Compat.H): amrex::Initialize / Finalize and the ParmParse table.
#include "AMReX_Compat.H"

#include <cstdlib>
#include <fstream>

namespace amrex {

void clear_comm_caches ();      // halo-plan caches (AMReX_MultiFab.cpp)

std::map<std::string, std::vector<std::string>>& ParmParse::table ()
{
    static std::map<std::string, std::vector<std::string>> t;
    return t;
}

// "name = v1 v2 ..." (the reference's inputs syntax, Src/Base/AMReX_ParmParse.cpp: '#' starts a comment, values are
// separated by blanks, double quotes group one value); a later definition replaces an earlier one
void ParmParse::addDefinition (std::string const& text)
{
    std::string line = text.substr(0, text.find('#'));
    const auto eq = line.find('=');
    if (eq == std::string::npos) { return; }
    auto trim = [] (std::string s) {
        const auto a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    const std::string name = trim(line.substr(0, eq));
    if (name.empty()) { return; }
    std::vector<std::string> vals;
    std::string rest = line.substr(eq + 1), cur;
    bool quoted = false;
    for (char c : rest) {
        if (c == '"') { if (quoted) { vals.push_back(cur); cur.clear(); } quoted = !quoted; continue; }
        if (!quoted && (c == ' ' || c == '\t' || c == '\r' || c == '\n')) { if (!cur.empty()) { vals.push_back(cur); cur.clear(); } continue; }
        cur.push_back(c);
    }
    if (!cur.empty()) { vals.push_back(cur); }
    table()[name] = vals;
}

void ParmParse::addFile (std::string const& path)
{
    std::ifstream f(path);
    if (!f.good()) { Abort("ParmParse: cannot open inputs file " + path); }
    std::string line;
    while (std::getline(f, line)) { addDefinition(line); }
}

void Initialize (int& argc, char**& argv)
{
    ParmParse::clear();
    int first = 1;
    if (argc > 1 && std::string(argv[1]).find('=') == std::string::npos) { ParmParse::addFile(argv[1]); first = 2; }
    for (int i = first; i < argc; ++i) { ParmParse::addDefinition(argv[i]); }
    const char* lr = std::getenv("LOCAL_RANK");
    Gpu::Initialize(lr ? std::atoi(lr) : 0);
}

void Finalize ()
{
    clear_comm_caches(); LevelLayout::clearCache(); ParallelDescriptor::FinalizeComm();
    Gpu::Finalize();
    ParmParse::clear();
}

bool TilingIfNotGPU () noexcept { return false; }

} // namespace amrex
