// vx3_node_worker -i <task.vxt> -o <report.vxr> [-f]
// Drop-in for the reference worker (src/Executables/vx3_node_worker.cu:32-141): same flags, same .vxt input, same
// .vxr report, .history text on stdout.  All work happens in vx3_worker_run_vxt (libvx3_b200.so).
#include <sys/stat.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "vx3_model.h"
#include "vx3_worker.h"

static void usage() {
    printf("This application is called by voxcraft-sim. If you need to use this directly, please refer to voxcraft-sim.\n"
           "  -h [ --help ]        produce help message\n"
           "  -i [ --input ] arg   Set input .vxt task file (base VXA, input dir, VXD list).\n"
           "  -o [ --output ] arg  Set output file path for report. (e.g. report_1.xml)\n"
           "  -f [ --force ]       Overwrite output file if exists.\n"
           "  -n [ --gpus ] arg    Use at most this many GPUs (default: all).\n\n");
}

int main(int argc, char **argv) {
    std::string input, output;
    bool force = false;
    int gpus = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-i" || a == "--input") && i + 1 < argc) input = argv[++i];
        else if ((a == "-o" || a == "--output") && i + 1 < argc) output = argv[++i];
        else if ((a == "-n" || a == "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "-f" || a == "--force") force = true;
        else { usage(); return 1; }
    }
    if (input.empty() || output.empty()) { usage(); return 1; }
    struct stat st;
    if (stat(output.c_str(), &st) == 0 && S_ISREG(st.st_mode) && !force) {
        printf("Error: output file exists.\n\n");
        usage();
        return 1;
    }
    if (stat(input.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) {
        printf("Error: input file not found.\n\n");
        usage();
        return 1;
    }
    vx3_worker_opts o;
    memset(&o, 0, sizeof(o));
    o.n_devices = gpus;
    o.emit_history = 1;
    o.verbose = 1;
    int rc = vx3_worker_run_vxt(input.c_str(), output.c_str(), &o);
    if (rc != 0) {
        const char *e1 = vx3_model_last_error(), *e2 = vx3_last_error();
        fprintf(stderr, "ERROR (%d): %s %s\n", rc, e1 ? e1 : "", e2 ? e2 : "");
        return 1;
    }
    return 0;
}
