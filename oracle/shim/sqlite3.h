/* TEST INFRASTRUCTURE ONLY (oracle/): minimal sqlite3 prototype header.
 * The image ships libsqlite3.so.0 (3.45.1) without its development header; this
 * declares exactly the public sqlite3 C API entry points the reference calls
 * (reference src/bam2db_ds.c:127-689) so the UNMODIFIED reference sources can be
 * compiled where they lie. Signatures are the documented public sqlite3 API. */
#ifndef FASTF_ORACLE_SQLITE3_SHIM_H
#define FASTF_ORACLE_SQLITE3_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct sqlite3 sqlite3;
typedef struct sqlite3_stmt sqlite3_stmt;
typedef long long sqlite3_int64;
#define SQLITE_OK 0
#define SQLITE_ROW 100
#define SQLITE_DONE 101
#define SQLITE_INTEGER 1
#define SQLITE_FLOAT 2
#define SQLITE_TEXT 3
#define SQLITE3_TEXT 3
#define SQLITE_BLOB 4
#define SQLITE_NULL 5
typedef void (*sqlite3_destructor_type)(void *);
#define SQLITE_STATIC ((sqlite3_destructor_type)0)
#define SQLITE_TRANSIENT ((sqlite3_destructor_type)-1)
int sqlite3_open(const char *filename, sqlite3 **ppDb);
int sqlite3_close(sqlite3 *);
int sqlite3_exec(sqlite3 *, const char *sql, int (*cb)(void *, int, char **, char **), void *, char **errmsg);
const char *sqlite3_errmsg(sqlite3 *);
void sqlite3_free(void *);
int sqlite3_prepare_v2(sqlite3 *db, const char *zSql, int nByte, sqlite3_stmt **ppStmt, const char **pzTail);
int sqlite3_bind_text(sqlite3_stmt *, int, const char *, int, void (*)(void *));
int sqlite3_bind_blob(sqlite3_stmt *, int, const void *, int n, void (*)(void *));
int sqlite3_bind_int(sqlite3_stmt *, int, int);
int sqlite3_bind_int64(sqlite3_stmt *, int, sqlite3_int64);
int sqlite3_bind_null(sqlite3_stmt *, int);
int sqlite3_step(sqlite3_stmt *);
int sqlite3_reset(sqlite3_stmt *);
int sqlite3_finalize(sqlite3_stmt *);
int sqlite3_column_count(sqlite3_stmt *);
int sqlite3_column_type(sqlite3_stmt *, int iCol);
int sqlite3_column_int(sqlite3_stmt *, int iCol);
sqlite3_int64 sqlite3_column_int64(sqlite3_stmt *, int iCol);
double sqlite3_column_double(sqlite3_stmt *, int iCol);
const unsigned char *sqlite3_column_text(sqlite3_stmt *, int iCol);
const void *sqlite3_column_blob(sqlite3_stmt *, int iCol);
int sqlite3_column_bytes(sqlite3_stmt *, int iCol);
const char *sqlite3_column_name(sqlite3_stmt *, int N);
#ifdef __cplusplus
}
#endif
#endif
