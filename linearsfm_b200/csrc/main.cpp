// LinearSFM executable: same command line as the reference's linux/src/LinearSFM/LinearSFM.cpp:9-18
// (constructs the implementation and calls run(argc, argv); always returns 0).
#include "../../include/linearsfm_b200.h"

int main(int argc, char *argv[])
{
    lsfm_cli_main(argc, argv);
    lsfm_shutdown();
    return 0;
}
