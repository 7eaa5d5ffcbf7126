// LinearSFM executable: same command line as the reference's linux/src/LinearSFM/LinearSFM.cpp:9-18
// (constructs the implementation and calls run(argc, argv)).  Usage problems exit 0 like the
// reference; a failed load / solve / write exits 1 so that scripts can detect it.
#include "../../include/linearsfm_b200.h"

int main(int argc, char *argv[])
{
    int rc = lsfm_cli_main(argc, argv);
    lsfm_shutdown();
    return rc;
}
