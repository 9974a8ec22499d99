// voxcraft-sim -i <dir with base.vxa + *.vxd> [-o report.xml] [-w worker] [-f] [-l]
// Drop-in for the reference front end (src/Executables/voxcraft-sim.cpp:25-132): validates the arguments, writes
// workspace/locally/<time>.<hash>.vxt, spawns "<worker> -i <vxt> -o <vxr>", copies the .vxr to the output.
#include <dirent.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <functional>
#include <string>
#include <vector>

static void usage() {
    printf("Thank you for using Voxelyze3 (B200 engine). This program should be run on a computer that has GPUs.\n"
           "Allowed options:\n"
           "  -h [ --help ]        produce help message\n"
           "  -l [ --locally ]     If this machine already has GPUs, locally run tasks on this machine.\n"
           "  -i [ --input ] arg   Set input directory path which contains a generation of VXA files.\n"
           "  -o [ --output ] arg  Set output file path for report. (e.g. report_1.xml)\n"
           "  -w [ --worker ] arg  Specify which worker you want to use. vx3_node_worker by default.\n"
           "  -f [ --force ]       Overwrite output file if exists.\n\n");
}
static bool is_file(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
static bool is_dir(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

int main(int argc, char **argv) {
    std::string input, output, worker = "./vx3_node_worker";
    bool force = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-i" || a == "--input") && i + 1 < argc) input = argv[++i];
        else if ((a == "-o" || a == "--output") && i + 1 < argc) output = argv[++i];
        else if ((a == "-w" || a == "--worker") && i + 1 < argc) worker = argv[++i];
        else if (a == "-f" || a == "--force") force = true;
        else if (a == "-l" || a == "--locally") {}
        else { usage(); return 1; }
    }
    if (input.empty()) { usage(); return 1; }
    printf("Running simulation locally by default.\n");
    if (!output.empty() && is_file(output) && !force) { printf("Error: output file exists.\n\n"); usage(); return 1; }
    if (!is_dir(input)) { printf("Error: input directory not found.\n\n"); usage(); return 1; }
    while (input.size() > 1 && input.back() == '/') input.pop_back();
    if (!is_file(input + "/base.vxa")) { printf("No base.vxa found in input directory.\n\n"); usage(); return 1; }
    if (!is_file(worker)) { printf("Need an executable worker but nothing found.\n\n"); usage(); return 1; }
    mkdir("workspace", 0777);
    mkdir("workspace/locally", 0777);
    std::vector<std::string> vxds;
    if (DIR *d = opendir(input.c_str())) {
        while (dirent *e = readdir(d)) {
            std::string n = e->d_name, low = n;
            std::transform(low.begin(), low.end(), low.begin(), ::tolower);
            if (low.size() > 4 && low.substr(low.size() - 4) == ".vxd") vxds.push_back(n);
        }
        closedir(d);
    }
    std::sort(vxds.begin(), vxds.end());
    char tbuf[32];
    time_t now = time(nullptr);
    strftime(tbuf, sizeof(tbuf), "%Y%m%d%H%M%S", localtime(&now));
    const std::string stem = std::string("workspace/locally/") + tbuf + "." + std::to_string(std::hash<std::string>{}(input));
    const std::string vxt = stem + ".vxt", vxr = stem + ".vxr";
    FILE *f = fopen(vxt.c_str(), "w");
    if (!f) { printf("ERROR: cannot write %s\n", vxt.c_str()); return 1; }
    fprintf(f, "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n<vxa>%s/base.vxa</vxa><input_dir>%s</input_dir><vxd>", input.c_str(), input.c_str());
    for (auto &v : vxds) fprintf(f, "<f>%s</f>", v.c_str());
    fprintf(f, "</vxd>\n");
    fclose(f);
    printf("%s -i %s -o %s\n", worker.c_str(), vxt.c_str(), vxr.c_str());
    fflush(stdout);
    pid_t pid = fork();
    if (pid == 0) {
        execl(worker.c_str(), worker.c_str(), "-i", vxt.c_str(), "-o", vxr.c_str(), "-f", (char *)nullptr);
        _exit(127);
    }
    int status = 0;
    waitpid(pid, &status, 0);
    if (is_file(vxr)) {
        if (!output.empty()) {
            std::string text;
            FILE *in = fopen(vxr.c_str(), "rb"), *out = fopen(output.c_str(), "wb");
            if (!in || !out) { printf("ERROR: Failed to copy result file: %s.\n", vxr.c_str()); return 1; }
            char buf[65536];
            size_t n;
            while ((n = fread(buf, 1, sizeof(buf), in)) > 0) fwrite(buf, 1, n, out);
            fclose(in);
            fclose(out);
        }
    } else
        printf("File not exist: %s. Worker failed to finish the job.\n", vxr.c_str());
    return 0;
}
