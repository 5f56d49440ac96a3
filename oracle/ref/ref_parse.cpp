// Feeds stdin to the REFERENCE's own output scrapers (ref common/parser.h:
// PETScOutputParser::parse1 :149-155, BoomerAMGParser::parse :181-266), compiled from
// where they lie, and prints what they extracted.  TEST INFRASTRUCTURE (oracle/): pins
// the text the dealii_compat layer prints for `output_details` and `-ksp_monitor`.
//
//   ref_parse boomeramg < text   -> rows / nze / sparsity / grid operator memory
//   ref_parse ksp < text         -> one residual per line (%.17e)
#include <cstdio>
#include <cstring>
#include <iterator>

#include "parser.h"  // the reference's file

int main(int argc, char** argv) {
  if (argc != 2) return 2;
  std::string text((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
  if (!std::strcmp(argv[1], "boomeramg")) {
    BoomerAMGParser p;
    if (!p.parse(text)) return 1;
    for (double v : p.get_rows()) std::printf("%.17e ", v);
    std::printf("\n");
    for (double v : p.get_nze()) std::printf("%.17e ", v);
    std::printf("\n");
    for (double v : p.get_sparsity()) std::printf("%.17e ", v);
    std::printf("\n%.17e %.17e %.17e\n", p.get_grid(), p.get_operator(), p.get_memory());
    return 0;
  }
  PETScOutputParser p(PETScOutputParser::monitor);
  p.parse(text);
  for (double v : p.get(PETScOutputParser::preconditioned_residual)) std::printf("%.17e\n", v);
  return 0;
}
