# gpu_solve.R — R host side of the B200 solve path (binds rpkg/src/r_glue.c through .Call).
#
# What a maintainer of benet1one/EasyLP changes (see INTEGRATION.md): the body of `$solve()` and of
# `private$feasible()` in R/class.R; DESCRIPTION drops `lpSolveAPI` from Imports; NAMESPACE gains
# `useDynLib(easylp, .registration = TRUE)`.  The six public methods keep their signatures.
#
# NOTE: R is not installed in the image this repository is built in; this file has been reviewed, not run.
# Its behaviour is mirrored, statement for statement, by easylp_b200/model.py (which the test-suite drives).

# lp_solve status codes -> the strings of the reference's switch (R/class.R:279-295), unchanged.
easylp_status_string <- function(status) {
    switch(
        as.character(status),
        "0" = "optimal",
        "1" = "sub-optimal",
        "2" = "unfeasible",
        "3" = "unbounded",
        "4" = "degenerate model",
        "5" = "numerical failure encountered",
        "6" = "process aborted",
        "7" = "timeout",
        "undocumented status"
    )
}

# Canonical CSR of a dense constraint matrix (entries != 0, row-major, ascending column).  Used while the DSL
# still stores `constraint$mat` densely; the sparse term-list DSL calls easylp_assemble_csr on its terms instead.
easylp_dense_to_csr <- function(mat) {
    nz <- which(t(mat) != 0, arr.ind = TRUE)          # row-major order of the original matrix
    list(row_ptr = c(0L, cumsum(tabulate(nz[, 2L], nbins = nrow(mat)))),
         col_idx = as.integer(nz[, 1L]),
         vals = t(mat)[nz])
}

# Device assembly of a term list (1-based row/col ids, emission order) into the canonical CSR:
# stable sort by (row, col), left-to-right fold of duplicates — the Reduce(`+`) of sum()/sum_for
# (R/methods.R:244-257) — zero drop, prefix sum, scatter.
easylp_assemble <- function(term_row, term_col, term_val, nrow, ncol) {
    .Call("easylp_assemble_csr", as.integer(term_row), as.integer(term_col), as.double(term_val),
          as.integer(nrow), as.integer(ncol))
}

# Replacement for the body of easylp$solve (R/class.R:251-302).  `self`/`private` are the R6 bindings.
easylp_solve_impl <- function(self, private, ...) {
    if (private$n_var == 0L)
        stop("Problem contains no variables.")
    if (all(self$objective_fun == 0))
        stop("Must specify objective function.")
    if (!is.element(private$dir, c("min", "max")))
        stop("Direction must be either 'min' or 'max'.")
    if (self$any_integer())
        stop("integer/binary variables need lp_solve's branch and bound; the B200 path solves LPs only")

    control <- list(...)
    known <- c("timeout", "epsilon", "verbose", "gpu.tol", "gpu.max_iter", "gpu.method", "gpu.transpose")
    for (k in setdiff(names(control), known))
        warning("lp.control option '", k, "' has no meaning on the GPU path and is ignored")
    control <- control[intersect(names(control), known)]

    csr <- if (!is.null(self$constraint$csr)) self$constraint$csr else easylp_dense_to_csr(self$constraint$mat)
    lower <- unlist(lapply(self$variables, function(x) rep(x$bound[1L], length(x$ind))))
    upper <- unlist(lapply(self$variables, function(x) rep(x$bound[2L], length(x$ind))))

    res <- .Call("easylp_solve_lp", as.integer(csr$row_ptr), as.integer(csr$col_idx), as.double(csr$vals),
                 as.character(self$constraint$dir), as.double(self$constraint$rhs), as.double(self$objective_fun),
                 private$dir == "max", as.double(lower), as.double(upper), control)

    private$objval <- res$objval |> large_to_infinity()
    private$sol[] <- res$x |> large_to_infinity()
    private$stat <- easylp_status_string(res$status)
    for (x in self$variables)  if (x$bound[1L] > x$bound[2L])
        private$stat <- "unfeasible"
    self$pointer <- res$stats            # was: the lpSolveAPI handle (R/class.R:300); now the per-solve statistics
    private$duals <- res$y
    invisible(self)
}

# Replacement for private$feasible (R/class.R:533-540): mat %*% sol + compare_tol on the device.
easylp_feasible_impl <- function(self, private, tol = 2e-8) {
    csr <- if (!is.null(self$constraint$csr)) self$constraint$csr else easylp_dense_to_csr(self$constraint$mat)
    stopifnot(length(csr$row_ptr) > 1L)
    nam <- self$constraint$rownames
    if (is.null(nam)) nam <- rownames(self$constraint$mat)
    nam[nam == ""] <- which(nam == "")
    sol <- private$sol
    sol[!is.finite(sol)] <- 0
    .Call("easylp_check_feasible", as.integer(csr$row_ptr), as.integer(csr$col_idx), as.double(csr$vals),
          as.double(sol), as.character(self$constraint$dir), as.double(self$constraint$rhs), tol) |>
        rlang::set_names(nam)
}

#' Solve a batch of small dense LPs on the GPU (additive entry point; BASELINE config 3).
#'
#' min/max c'x  s.t.  A x (dir) b,  lower <= x <= upper, for every slice of the batch.
#' @param A numeric array with dim c(m, n, B).
#' @param b numeric matrix m x B.   @param c numeric matrix n x B.
#' @param lower,upper numeric matrices n x B (or NULL for 0 / +Inf).
#' @param dir character(m) of "<=", ">=", "==" (recycled over the batch) or NULL for all "<=".
#' @export
easylp_solve_batch <- function(A, b, c, lower = NULL, upper = NULL, dir = NULL, sense = c("min", "max"), ...) {
    sense <- match.arg(sense)
    stopifnot(length(dim(A)) == 3L)
    At <- aperm(A, c(2L, 1L, 3L))         # n x m x B: every LP's rows contiguous for the [B][m][n] ABI layout
    res <- .Call("easylp_solve_batch", At, b, c, lower, upper, dir, sense == "max", list(...))
    res$status <- vapply(res$status, easylp_status_string, "")
    res
}
