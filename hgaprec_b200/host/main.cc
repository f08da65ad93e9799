// main.cc -- the `hgaprec` command line over the B200 engine.
//
// Same flag surface and dispatch as the reference's src/main.cc:99-232, 342-361:
//   hgaprec -dir D -n N -m M -k K [-hier] [-bias] [-binary-data] [-novb]
//           [-rfreq R] [-max-iterations T] [-seed S] [-label L]
//           [-rating-threshold V] [-gen-ranking]
// Flags the reference parses but never reads on this path (-a -b -c -d -load
// -online ...) are accepted with the same arity and ignored the same way; its
// other-model baselines (-nmf -lda -chi ...: external programs, SURVEY.md 2 rows
// 12-18) are outside this path and rejected.  Extensions: -device G picks the
// CUDA device, -gpus N shards the users over GPUs 0..N-1 of this box (one process,
// hpf_config.n_devices), -csr-cache keeps the parsed training matrix beside train.tsv.
#include <assert.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <string>

#include "hgaprec.hh"

using namespace hpfhost;

static volatile sig_atomic_t g_save_state_now = 0;
static void term_handler(int) { g_save_state_now = 1; } // src/main.cc:19-30

static void plog(FILE *f, const char *k, double v) { fprintf(f, "%s: %.9f\n", k, v); }
static void plog(FILE *f, const char *k, unsigned v) { fprintf(f, "%s: %d\n", k, v); }
static void plog(FILE *f, const char *k, bool v) { fprintf(f, "%s: %s\n", k, v ? "True" : "False"); }

int main(int argc, char **argv)
{
  signal(SIGTERM, term_handler);
  if (argc <= 1) {
    printf("gaprec -dir <netflix-dataset-dir> -n <users>-m <movies> -k <dims> -label <out-dir-tag>\n");
    exit(0);
  }
  Options o;
  o.save_state_now = &g_save_state_now;
  for (int i = 1; i < argc; ++i) {
    const char *a = argv[i];
    #define NEXT (i + 1 < argc ? argv[++i] : "")
    if (!strcmp(a, "-dir")) { o.dir = NEXT; fprintf(stdout, "+ dir = %s\n", o.dir.c_str()); }
    else if (!strcmp(a, "-n")) { o.n = atoi(NEXT); fprintf(stdout, "+ n = %d\n", o.n); }
    else if (!strcmp(a, "-m")) { o.m = atoi(NEXT); fprintf(stdout, "+ m = %d\n", o.m); }
    else if (!strcmp(a, "-k")) { o.k = atoi(NEXT); fprintf(stdout, "+ k = %d\n", o.k); }
    else if (!strcmp(a, "-rfreq")) { o.rfreq = atoi(NEXT); fprintf(stdout, "+ rfreq = %d\n", o.rfreq); }
    else if (!strcmp(a, "-label")) o.label = NEXT;
    else if (!strcmp(a, "-logl")) o.logl = true;
    else if (!strcmp(a, "-max-iterations")) o.max_iterations = atoi(NEXT);
    else if (!strcmp(a, "-seed")) o.seed = atof(NEXT);
    else if (!strcmp(a, "-a")) o.a = atof(NEXT);
    else if (!strcmp(a, "-b")) o.b = atof(NEXT);
    else if (!strcmp(a, "-c")) o.c = atof(NEXT);
    else if (!strcmp(a, "-d")) o.d = atof(NEXT);
    else if (!strcmp(a, "-binary-data")) o.binary_data = true;
    else if (!strcmp(a, "-bias")) o.bias = true;
    else if (!strcmp(a, "-hier")) o.hier = true;
    else if (!strcmp(a, "-novb")) o.vb = false;
    else if (!strcmp(a, "-gen-ranking")) o.gen_ranking = true;
    else if (!strcmp(a, "-rating-threshold")) o.rating_threshold = atoi(NEXT);
    else if (!strcmp(a, "-device")) o.device = atoi(NEXT);
    else if (!strcmp(a, "-gpus")) o.gpus = atoi(NEXT);
    else if (!strcmp(a, "-csr-cache")) o.csr_cache = true;
    else if (!strcmp(a, "-load") || !strcmp(a, "-nmi") || !strcmp(a, "-wals_l") || !strcmp(a, "-wals_C")) (void)NEXT; // parsed, unused
    else if (!strcmp(a, "-batch") || !strcmp(a, "-p") || !strcmp(a, "-strid") || !strcmp(a, "-gen-heldout") ||
             !strcmp(a, "-pred-accuracy") || !strcmp(a, "-gt-accuracy") || !strcmp(a, "-netflix") || !strcmp(a, "-mendeley") ||
             !strcmp(a, "-movielens") || !strcmp(a, "-echonest")) {} // parsed, no effect on this path
    else if (!strcmp(a, "-online")) { printf("Quitting. Online inference not implemented.\n"); exit(0); }
    else if (!strcmp(a, "-nmf") || !strcmp(a, "-lda") || !strcmp(a, "-vwlda") || !strcmp(a, "-chi") || !strcmp(a, "-ctr") ||
             !strcmp(a, "-mle-user") || !strcmp(a, "-mle-item") || !strcmp(a, "-canny") || !strcmp(a, "-nyt") ||
             !strcmp(a, "-msr") || !strcmp(a, "-rmse") || !strcmp(a, "-test") || !strcmp(a, "-write-training") ||
             !strcmp(a, "-nmfload") || !strcmp(a, "-vwload") || !strcmp(a, "-als") || !strcmp(a, "-wals") ||
             !strcmp(a, "-chinmf") || !strcmp(a, "-climf")) {
      fprintf(stdout, "error: option %s selects a code path outside the B200 engine (see DESIGN.md, out of scope)\n", a);
      fflush(stdout);
      exit(2);
    } else {
      fprintf(stdout, "error: unknown option %s\n", a); // src/main.cc:226-229
      fflush(stdout);
      assert(0);
      abort();
    }
    #undef NEXT
  }
  if (o.k == 0 || o.n == 0 || o.m == 0 || o.dir.empty()) {
    fprintf(stderr, "hgaprec: -dir, -n, -m and -k are required\n");
    return -1;
  }
  o.prefix = o.make_prefix();
  fprintf(stdout, "+ Creating directory %s\n", o.prefix.c_str());
  fflush(stdout);
  mkdir(o.prefix.c_str(), 0777);
  FILE *pl = fopen((o.prefix + "/param.txt").c_str(), "w");
  if (!pl) {
    printf("cannot open param file\n");
    exit(-1);
  }
  // Env::Env's plog block (src/env.hh:386-405)
  plog(pl, "n", o.n); plog(pl, "k", o.k); plog(pl, "t", 2u);
  plog(pl, "test_ratio", 0.2); plog(pl, "validation_ratio", 0.01); plog(pl, "seed", o.seed);
  plog(pl, "a", o.a); plog(pl, "b", o.b); plog(pl, "c", o.c); plog(pl, "d", o.d);
  plog(pl, "reportfreq", o.rfreq); plog(pl, "vb", o.vb); plog(pl, "bias", o.bias); plog(pl, "hier", o.hier);
  fflush(pl);

  Ratings ratings(o.n, o.m, o.binary_data, o.rating_threshold);
  fprintf(stdout, "+ reading ratings dataset from %s\n", o.dir.c_str());
  fflush(stdout);
  std::string err;
  if (!ratings.read_train(o.dir, &err, o.csr_cache)) {
    fprintf(stderr, "error: %s\n", err.c_str());
    exit(-1);
  }
  if (o.csr_cache) fprintf(stdout, "+ csr cache: %s\n", ratings.cache_note().c_str());
  ratings.write_marginals(o.prefix);
  fprintf(pl, "training ratings: %d\n", (int)ratings.nratings());
  fprintf(pl, "statistics: read %d users, %d movies, %d ratings\n", ratings.n(), ratings.m(), (int)ratings.nratings());
  fclose(pl);
  if (ratings.n() == 0 || ratings.m() == 0) {
    fprintf(stderr, "error reading dataset from dir %s; quitting\n", o.dir.c_str());
    return -1;
  }

  HGAPRec hgaprec(o, ratings);
  if (o.gen_ranking) {
    hgaprec.gen_ranking_for_users(true);
    exit(0);
  }
  if (o.bias && !o.hier) hgaprec.vb_bias();
  else if (o.hier) hgaprec.vb_hier();
  else hgaprec.vb();
  return 0;
}
