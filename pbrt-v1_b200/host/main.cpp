// main.cpp -- the `pbrt` executable (renderer/pbrt.cpp:28-51): banner, init,
// parse each file named on the command line (or stdin), cleanup.
#include "pbrt.h"
#include "api.h"

int main(int argc, char *argv[]) {
  printf("pbrt version %1.3f (B200 render hot path)\n", PBRT_VERSION);
  fflush(stdout);
  pbrtInit();
  if (argc == 1) {
    ParseFile("-");
  } else {
    for (int i = 1; i < argc; i++)
      if (!ParseFile(argv[i])) Error("Couldn't open scene file \"%s\"\n", argv[i]);
  }
  pbrtCleanup();
  return 0;
}
