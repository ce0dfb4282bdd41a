// Host-only unit test of amrex::ParmParse (compat layer): inputs-file syntax, command-line overrides, type conversion.
#include <AMReX_ParmParse.H>
#include <cstdio>
#include <fstream>
using namespace amrex;

#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main (int argc, char* argv[])
{
    const char* path = argv[1];
    {
        std::ofstream f(path);
        f << "# comment line\n"
          << "max_level = 1\n"
          << "n_cell = 128   # trailing comment\n"
          << "composite_solve = 0\n"
          << "agglomeration=true\n"
          << "tol = 1.e-10\n"
          << "name = \"two words\" second\n"
          << "n_cell = 64\n"                    // a later definition wins
          << "amr.ref_ratio = 2 4\n"
          << "big = 1e3\n";
    }
    ParmParse::clear();
    ParmParse::addFile(path);
    ParmParse::addDefinition("prob_type=2");
    ParmParse::addDefinition("n_cell = 32");   // the command line overrides the file
    ParmParse pp;
    int max_level = -1, n_cell = -1, prob_type = -1, missing = 7, big = 0;
    bool composite = true, agg = false;
    double tol = 0; std::string name;
    CHECK(pp.query("max_level", max_level) == 1 && max_level == 1);
    CHECK(pp.query("n_cell", n_cell) == 1 && n_cell == 32);
    CHECK(pp.query("prob_type", prob_type) == 1 && prob_type == 2);
    CHECK(pp.query("not_there", missing) == 0 && missing == 7);
    CHECK(pp.query("composite_solve", composite) == 1 && composite == false);
    CHECK(pp.query("agglomeration", agg) == 1 && agg == true);
    CHECK(pp.query("tol", tol) == 1 && tol == 1.e-10);
    CHECK(pp.query("big", big) == 1 && big == 1000);
    CHECK(pp.query("name", name) == 1 && name == "two words");
    CHECK(pp.countval("name") == 2 && pp.contains("tol") && !pp.contains("nope"));
    ParmParse amr("amr");
    std::vector<int> rr;
    CHECK(amr.queryarr("ref_ratio", rr) == 1 && rr.size() == 2 && rr[0] == 2 && rr[1] == 4);
    pp.add("added", 3.5);
    double added = 0; CHECK(pp.query("added", added) == 1 && added == 3.5);
    std::printf("PARMPARSE OK\n");
    return 0;
}
