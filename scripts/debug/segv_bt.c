// Debug aid: print a native backtrace on SIGSEGV (module+offset lines; resolve with addr2line -e <lib> <offset>).
// Build: gcc -shared -fPIC -o segv_bt.so segv_bt.c ; load with ctypes.CDLL after faulthandler.enable(), then call segv_bt_install().
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

static void on_segv(int sig, siginfo_t *si, void *uc)
{
	(void)uc;
	char msg[128];
	int n = snprintf(msg, sizeof msg, "\n[segv_bt] signal %d at address %p\n", sig, si->si_addr);
	if (write(2, msg, n) < 0) {}
	void *frames[64];
	int k = backtrace(frames, 64);
	backtrace_symbols_fd(frames, k, 2);
	signal(sig, SIG_DFL);
	raise(sig);
}

void segv_bt_install(void)
{
	static char stack[1 << 16];
	stack_t ss; ss.ss_sp = stack; ss.ss_size = sizeof stack; ss.ss_flags = 0;
	sigaltstack(&ss, 0);
	struct sigaction sa;
	memset(&sa, 0, sizeof sa);
	sa.sa_sigaction = on_segv;
	sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
	sigaction(SIGSEGV, &sa, 0);
	sigaction(SIGBUS, &sa, 0);
}
