/* Prototypes of the few sqlite3 entry points the host writer uses.  This image ships libsqlite3.so.0 (3.45) without its
 * header; these declarations follow the documented public C API (https://sqlite.org/c3ref) and link with -l:libsqlite3.so.0. */
#ifndef FASTF_SQLITE3_DECL_H
#define FASTF_SQLITE3_DECL_H
typedef struct sqlite3 sqlite3;
typedef struct sqlite3_stmt sqlite3_stmt;
#define SQLITE_OK 0
#define SQLITE_DONE 101
#define SQLITE_STATIC ((void (*)(void *))0)
#define SQLITE_TRANSIENT ((void (*)(void *)) - 1)
int sqlite3_open(const char *filename, sqlite3 **db);
int sqlite3_close(sqlite3 *db);
int sqlite3_exec(sqlite3 *db, const char *sql, int (*cb)(void *, int, char **, char **), void *arg, char **errmsg);
const char *sqlite3_errmsg(sqlite3 *db);
void sqlite3_free(void *p);
int sqlite3_prepare_v2(sqlite3 *db, const char *sql, int nbyte, sqlite3_stmt **stmt, const char **tail);
int sqlite3_bind_text(sqlite3_stmt *s, int i, const char *v, int n, void (*d)(void *));
int sqlite3_bind_blob(sqlite3_stmt *s, int i, const void *v, int n, void (*d)(void *));
int sqlite3_bind_int(sqlite3_stmt *s, int i, int v);
int sqlite3_bind_null(sqlite3_stmt *s, int i);
int sqlite3_step(sqlite3_stmt *s);
int sqlite3_reset(sqlite3_stmt *s);
int sqlite3_finalize(sqlite3_stmt *s);
int sqlite3_column_int(sqlite3_stmt *s, int col);
#define SQLITE_ROW 100
#endif
