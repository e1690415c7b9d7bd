/*
 * l9_main.c -- TEST INFRASTRUCTURE.  The reference's main() only calls the L5
 * handler (src/main.c:477-481); its L9 handler (main.c:362) is reachable by
 * editing the source (readme step 2).  Instead we compile main.c with
 * -Dmain=ref_main and call the non-static L9_data_handler() from here.
 * The handler keeps ~80 MB of frames + SLAM_attr on the stack at 16x1800
 * (SURVEY D6), so raise the stack limit and re-exec once.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <unistd.h>

void L9_data_handler(void);
void L5_IMU_data_handler(void);

int main(int argc, char **argv) {
    struct rlimit rl;
    if (getrlimit(RLIMIT_STACK, &rl) == 0 && rl.rlim_cur != RLIM_INFINITY &&
        rl.rlim_cur < (rlim_t)1 << 33 && !getenv("NAVSLAM_L9_REEXEC")) {
        rl.rlim_cur = rl.rlim_max;
        if (setrlimit(RLIMIT_STACK, &rl) == 0) {
            setenv("NAVSLAM_L9_REEXEC", "1", 1);
            execv("/proc/self/exe", argv);
        }
    }
    if (argc > 1 && strcmp(argv[1], "l5") == 0)
        L5_IMU_data_handler();
    else
        L9_data_handler();
    return 0;
}
