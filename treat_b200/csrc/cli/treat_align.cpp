// treat_align -- `treat align` (cmd/treat/main.go:91-138 flags, cmd/treat/align.go:59-122) on
// top of libtreat_b200: same options, byte-identical stdout, alignment done on the GPU.
//   treat_align [-t templates.fa] [-f reads.fa] [-1 S1 -2 S2] [-b T] [--offset N] [--device D]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../host/treat_host.hpp"

int main(int argc, char **argv)
{
    std::string t, f, s1, s2, base = "T";
    int offset = 0, device = 0;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        const char *v = i + 1 < argc ? argv[i + 1] : nullptr;
        if ((a == "-t" || a == "--template") && v) t = v, i++;
        else if ((a == "-f" || a == "--fragment") && v) f = v, i++;
        else if ((a == "-1" || a == "--s1") && v) s1 = v, i++;
        else if ((a == "-2" || a == "--s2") && v) s2 = v, i++;
        else if ((a == "-b" || a == "--base") && v) base = v, i++;
        else if (a == "--offset" && v) offset = atoi(v), i++;
        else if (a == "--device" && v) device = atoi(v), i++;
        else {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    treat_b200_ctx *ctx = nullptr;
    int rc = treat_b200_create(device, &ctx);
    if (rc != TREAT_B200_OK) {
        fprintf(stderr, "FATA %s\n", treat_b200_strerror(rc));
        return 1;
    }
    std::string err;
    const bool ok = treat::CliAlignStream(ctx, t, f, s1, s2, base, offset,
                                          [](const char *p, size_t n) { return fwrite(p, 1, n, stdout) == n; }, &err);
    if (!ok) fprintf(stderr, "FATA %s\n", err.c_str());  // logrus.Fatal
    treat_b200_destroy(ctx);
    return ok ? 0 : 1;
}
