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
        "9" = "the model was solved by presolve",
        "10" = "the branch and bound routine failed",
        "11" = "the branch and bound was stopped because of a break-at-first or break-at-value",
        "12" = "a feasible branch and bound solution was found",
        "13" = "no feasible branch and bound solution was found",
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

# Lowered assembly (include/easylp_abi.h "(1b)"): explicit terms plus the FAMILIES that for_split()/sum_for() record
# when their body, evaluated once with symbolic loop variables, is affine in indexed variables (the executable
# specification of that trace is easylp_b200/lower.py; see INTEGRATION.md).  `families` / `groups` are the lists
# rpkg/src/r_glue.c::pack_families documents.
easylp_assemble_lowered <- function(term_row, term_col, term_val, families, groups, nrow, ncol) {
    .Call("easylp_assemble_lowered", as.integer(term_row), as.integer(term_col), as.double(term_val),
          families, groups, as.integer(nrow), as.integer(ncol))
}

# R producer of `private$terms` — the term list the device-resident model is assembled from.  With the reference's
# dense DSL unchanged, `$con()` calls this right after `self$constraint <- join_constraints(...)` (R/class.R:206,217)
# with the rows it has just appended: their non-zero entries are added to the list in row-major order (the emission
# order of a dense row: one term per (row, col), so the fold on the device has nothing to add up), and the device
# model is dropped so that the next `$solve()` re-assembles it.  `$uncon()` / `$var()` call easylp_terms_reset() and
# the list is rebuilt from `self$constraint$mat` on demand.  The sparse term-list DSL (easylp_b200/model.py is its
# executable specification) appends its atoms' terms here instead and never materialises `constraint$mat`.
easylp_terms_append <- function(private, new_rows, first_row) {
    if (is.null(private$terms))
        private$terms <- list(row = integer(), col = integer(), val = double(), families = list(), groups = list())
    if (nrow(new_rows) > 0L) {
        tm <- t(new_rows)
        nz <- which(tm != 0, arr.ind = TRUE)               # column-major walk of the transpose = row-major walk of the rows
        private$terms$row <- c(private$terms$row, as.integer(first_row - 1L + nz[, 2L]))
        private$terms$col <- c(private$terms$col, as.integer(nz[, 1L]))
        private$terms$val <- c(private$terms$val, tm[nz])
    }
    private$model <- NULL
    invisible(private$terms)
}
easylp_terms_reset <- function(private) {
    private$terms <- NULL
    private$model <- NULL
    invisible(NULL)
}
easylp_terms_ensure <- function(self, private) {
    if (is.null(private$terms))
        easylp_terms_append(private, self$constraint$mat, 1L)
    private$terms
}

# Device-resident model: the same assembly, but the canonical CSR stays in HBM behind an external pointer (with a
# finalizer).  `$con()`/`$uncon()`/`$var()` set `private$model <- NULL`; `$solve()` builds it on demand and solves
# on it, so the matrix crosses PCIe at most once (as descriptors or terms) between `$con()` and the solution.
# A cloned or readRDS-restored object holds a NULL pointer: the glue raises, and easylp_model() rebuilds.
easylp_model <- function(self, private) {
    # rebuild only when there is no live device model: never built, invalidated by $con()/$var(), or a pointer that did
    # not survive a clone / readRDS (external pointers come back NULL).  Interrupts and real errors propagate.
    if (is.null(private$model) || !.Call("easylp_model_valid", private$model)) {
        t <- easylp_terms_ensure(self, private)     # list(row, col, val, families, groups): appended by $con(), see above
        private$model <- .Call("easylp_model_assemble", as.integer(t$row), as.integer(t$col), as.double(t$val),
                               t$families, t$groups, length(self$constraint$rhs), private$n_var)
    }
    private$model
}
easylp_model_csr <- function(model) .Call("easylp_model_csr", model)     # constraint$mat for printing / inspection

# Replacement for the body of easylp$solve (R/class.R:251-302).  `self`/`private` are the R6 bindings.
easylp_solve_impl <- function(self, private, ...) {
    if (private$n_var == 0L)
        stop("Problem contains no variables.")
    if (all(self$objective_fun == 0))
        stop("Must specify objective function.")
    if (!is.element(private$dir, c("min", "max")))
        stop("Direction must be either 'min' or 'max'.")

    control <- list(...)
    known <- c("timeout", "verbose", "gpu.tol", "gpu.max_iter", "gpu.method", "gpu.transpose", "gpu.devices")
    if ("epsilon" %in% names(control))      # lp_solve's integer-rounding tolerance: no counterpart here (use gpu.tol)
        warning("lp.control option 'epsilon' is lp_solve's integer-rounding tolerance and is ignored; the GPU path's optimality tolerance is gpu.tol")
    control$epsilon <- NULL
    for (k in setdiff(names(control), known))
        warning("lp.control option '", k, "' has no meaning on the GPU path and is ignored")
    control <- control[intersect(names(control), known)]

    lower <- unlist(lapply(self$variables, function(x) rep(x$bound[1L], length(x$ind))))
    upper <- unlist(lapply(self$variables, function(x) rep(x$bound[2L], length(x$ind))))

    res <- if (self$any_integer()) {
        # set.type(prob, columns, type) + lp_solve's branch and bound (R/class.R:264-276): the frontiers of the tree go
        # through the batched simplex kernel (easylp_b200/csrc/mip.cu).  Binary columns carry the bounds [0, 1].
        is_int <- unlist(lapply(self$variables, function(x) rep(x$integer || x$binary, length(x$ind))))
        csr <- if (!is.null(private$terms)) easylp_model_csr(easylp_model(self, private)) else easylp_dense_to_csr(self$constraint$mat)
        .Call("easylp_solve_mip", as.integer(csr$row_ptr), as.integer(csr$col_idx), as.double(csr$vals),
              as.character(self$constraint$dir), as.double(self$constraint$rhs), as.double(self$objective_fun),
              private$dir == "max", as.double(lower), as.double(upper), as.logical(is_int), control)
    } else if (!is.null(private$terms)) {
        # sparse term-list DSL: solve on the device-resident matrix (rebuilt if the pointer did not survive a clone)
        solve_on <- function() .Call("easylp_model_solve", easylp_model(self, private), as.character(self$constraint$dir),
                                     as.double(self$constraint$rhs), as.double(self$objective_fun), private$dir == "max",
                                     as.double(lower), as.double(upper), control)
        solve_on()
    } else {
        # smallest drop-in: the reference's dense constraint$mat, converted on the host
        csr <- easylp_dense_to_csr(self$constraint$mat)
        .Call("easylp_solve_lp", as.integer(csr$row_ptr), as.integer(csr$col_idx), as.double(csr$vals),
              as.character(self$constraint$dir), as.double(self$constraint$rhs), as.double(self$objective_fun),
              private$dir == "max", as.double(lower), as.double(upper), control)
    }

    private$objval <- res$objval |> large_to_infinity()
    private$sol[] <- res$x |> large_to_infinity()
    private$stat <- easylp_status_string(res$status)
    for (x in self$variables)  if (x$bound[1L] > x$bound[2L])
        private$stat <- "unfeasible"
    self$pointer <- res$stats            # was: the lpSolveAPI handle (R/class.R:300); now the per-solve statistics
    private$duals <- res$y
    invisible(self)
}

# Replacement for the bodies of the active bindings sensitivity_objective / sensitivity_rhs (R/class.R:613-646): the
# get.sensitivity.obj / get.sensitivity.rhs calls become one .Call; the arrays keep their shape and dimnames.
easylp_sensitivity_impl <- function(self, private) {
    if (private$stat != "optimal")
        stop("Problem is not optimal.", call. = FALSE)
    if (self$any_integer())
        stop("Sensitivity unavailable for problems with integer/binary variables")
    lower <- unlist(lapply(self$variables, function(x) rep(x$bound[1L], length(x$ind))))
    upper <- unlist(lapply(self$variables, function(x) rep(x$bound[2L], length(x$ind))))
    csr <- if (!is.null(private$terms)) easylp_model_csr(easylp_model(self, private)) else easylp_dense_to_csr(self$constraint$mat)
    .Call("easylp_sensitivity", as.integer(csr$row_ptr), as.integer(csr$col_idx), as.double(csr$vals),
          as.character(self$constraint$dir), as.double(self$constraint$rhs), as.double(self$objective_fun),
          private$dir == "max", as.double(lower), as.double(upper), list())
}
easylp_sensitivity_objective_impl <- function(self, private) {
    sens <- easylp_sensitivity_impl(self, private)
    objective <- array(dim = c(length(self$objective_fun), 3L),
                       dimnames = list(Variable = names2(private$sol), Bound = c("Lower", "Current", "Upper")))
    objective[, "Lower"] <- sens$objfrom
    objective[, "Upper"] <- sens$objtill
    objective[, "Current"] <- self$objective_fun
    objective
}
easylp_sensitivity_rhs_impl <- function(self, private) {
    sens <- easylp_sensitivity_impl(self, private)
    rhs <- array(dim = c(length(self$constraint$rhs), 3L),
                 dimnames = list(Constraint = self$constraint$rownames, Bound = c("Lower", "Current", "Upper")))
    rhs[, "Lower"] <- sens$rhsfrom
    rhs[, "Upper"] <- sens$rhstill
    rhs[, "Current"] <- self$constraint$rhs
    rhs
}

# Replacement for private$feasible (R/class.R:533-540): mat %*% sol + compare_tol on the device.
easylp_feasible_impl <- function(self, private, tol = 2e-8) {
    csr <- if (!is.null(private$terms)) easylp_model_csr(easylp_model(self, private)) else easylp_dense_to_csr(self$constraint$mat)
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
