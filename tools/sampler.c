/* LD_PRELOAD sampling profiler (development tool): SIGPROF every 1 ms of process CPU time, records the two innermost
 * return addresses inside libmm2b200.so / the executable, dumps "count addr" lines at exit for addr2line. */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <dlfcn.h>
#define MAXS (1 << 20)
static void *samp[MAXS][10];
static volatile int n_samp;
static void on_prof(int sig)
{
	void *bt[12];
	int k, n = backtrace(bt, 12), i = __sync_fetch_and_add(&n_samp, 1);
	if (i < MAXS) for (k = 0; k < 10; ++k) samp[i][k] = n > k + 2 ? bt[k + 2] : 0;
}
static void dump(void)
{
	FILE *fp = fopen(getenv("SAMPLER_OUT") ? getenv("SAMPLER_OUT") : "sampler.out", "w");
	int i, n = n_samp < MAXS ? n_samp : MAXS;
	struct itimerval z; memset(&z, 0, sizeof(z)); setitimer(ITIMER_PROF, &z, 0);
	for (i = 0; i < n; ++i) {
		int k;
		for (k = 0; k < 10; ++k) {
			Dl_info di;
			if (samp[i][k] && dladdr(samp[i][k], &di) && di.dli_fname)
				fprintf(fp, "%s+%lx%c", strrchr(di.dli_fname, '/') ? strrchr(di.dli_fname, '/') + 1 : di.dli_fname,
						(unsigned long)((char*)samp[i][k] - (char*)di.dli_fbase), k == 9 ? '\n' : '\t');
			else fprintf(fp, "?%c", k == 9 ? '\n' : '\t');
		}
	}
	fclose(fp);
}
__attribute__((constructor)) static void init(void)
{
	struct sigaction sa; struct itimerval it; void *bt[4];
	backtrace(bt, 4);
	memset(&sa, 0, sizeof(sa)); sa.sa_handler = on_prof; sa.sa_flags = SA_RESTART; sigaction(SIGPROF, &sa, 0);
	it.it_interval.tv_sec = 0, it.it_interval.tv_usec = 1000, it.it_value = it.it_interval;
	setitimer(ITIMER_PROF, &it, 0);
	atexit(dump);
}
