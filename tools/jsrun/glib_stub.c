/* Stub for the 16 glib entry points libQt6Core imports for its optional glib event
 * dispatcher.  The JS runner never starts an event loop (and sets QT_NO_GLIB=1), so none is
 * ever called; they exist only so that the dynamic loader can resolve libQt6Core. */
#include <stdlib.h>
#define STUB(n) void* n(void) { abort(); return 0; }
STUB(g_main_context_default) STUB(g_main_context_iteration) STUB(g_main_context_new)
STUB(g_main_context_pop_thread_default) STUB(g_main_context_push_thread_default)
STUB(g_main_context_ref) STUB(g_main_context_unref) STUB(g_main_context_wakeup)
STUB(g_source_add_poll) STUB(g_source_attach) STUB(g_source_destroy) STUB(g_source_new)
STUB(g_source_remove_poll) STUB(g_source_set_can_recurse) STUB(g_source_set_name) STUB(g_source_unref)
