#include "gle.hpp"

#include "exp_cr.hpp"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>

namespace acfhost {

namespace {
// constants of ACFcalculator.cpp:29-35
const double kSegMin = 0.2;
const double kSStart = 0.000001;
const double kSEnd = 1.0;
const double kSIncrement = 0.0001;
const double kFitWidth1 = 1000.0;
const double kFitWidth2 = 1000.0;

inline int sgn(double v) { return (v > 0) - (v < 0); }
}  // namespace

// F(s) = sum_i exp(-s*i*dt) * series[i] * dt      (ACFcalculator.cpp:77-87, same evaluation order).
// exp is the correctly rounded one (exp_cr.hpp): that is what the shipped reference binary's libm computes, and the
// segment search below turns a one-ulp difference here into a different fitting window.
double laplace(const double* series, double deltat, double s, int length)
{
    double F = 0.0;
    const int kBlock = 256;  // arguments and exponentials of a block at a time (exp_cr_array: four lanes where it can)
    double arg[kBlock], e[kBlock];
    for (int i0 = 0; i0 < length; i0 += kBlock) {
        const int m = length - i0 < kBlock ? length - i0 : kBlock;
        for (int j = 0; j < m; ++j) arg[j] = -s * (i0 + j) * deltat;
        exp_cr_array(arg, e, (size_t)m);
        for (int j = 0; j < m; ++j) F += e[j] * series[i0 + j] * deltat;
    }
    return F;
}

// ---------------------------------------------------------------------------------------------------
// Brent's method for a minimum (golden section + parabolic steps), in the variant used by Boost.Math:
// the search starts from the upper end of the interval, the golden ratio is the single-precision
// constant 0.3819660f, and the precision request is capped at half the mantissa.
// ---------------------------------------------------------------------------------------------------
std::pair<double, double> minimise_brent(const std::function<double(double)>& f, double lo, double hi, int bits)
{
    if (bits > DBL_MANT_DIG / 2) bits = DBL_MANT_DIG / 2;
    const double tolerance = ldexp(1.0, 1 - bits);
    const double golden = 0.3819660f;
    double x = hi, w = hi, v = hi, u;
    double fx = f(x), fw = fx, fv = fx, fu;
    double step = 0.0, prev_step = 0.0;
    for (std::uintmax_t iter = UINTMAX_MAX; iter != 0; --iter) {
        const double mid = (lo + hi) / 2;
        const double tol1 = tolerance * fabs(x) + tolerance / 4;
        const double tol2 = 2 * tol1;
        if (fabs(x - mid) <= tol2 - (hi - lo) / 2) break;
        bool golden_step = true;
        if (fabs(prev_step) > tol1) {
            // parabola through (v,fv), (w,fw), (x,fx)
            double r = (x - w) * (fx - fv);
            double q = (x - v) * (fx - fw);
            double p = (x - v) * q - (x - w) * r;
            q = 2 * (q - r);
            if (q > 0) p = -p;
            q = fabs(q);
            const double before_last = prev_step;
            prev_step = step;
            if (!(fabs(p) >= fabs(q * before_last / 2) || p <= q * (lo - x) || p >= q * (hi - x))) {
                step = p / q;
                u = x + step;
                if ((u - lo) < tol2 || (hi - u) < tol2) step = (mid - x) < 0 ? -fabs(tol1) : fabs(tol1);
                golden_step = false;
            }
        }
        if (golden_step) {
            prev_step = (x >= mid) ? lo - x : hi - x;
            step = golden * prev_step;
        }
        u = (fabs(step) >= tol1) ? x + step : (step > 0 ? x + fabs(tol1) : x - fabs(tol1));
        fu = f(u);
        if (fu <= fx) {
            if (u >= x)
                lo = x;
            else
                hi = x;
            v = w;
            w = x;
            x = u;
            fv = fw;
            fw = fx;
            fx = fu;
        } else {
            if (u < x)
                lo = u;
            else
                hi = u;
            if (fu <= fw || w == x) {
                v = w;
                w = u;
                fv = fw;
                fw = fu;
            } else if (fu <= fv || v == x || v == w) {
                v = u;
                fv = fu;
            }
        }
    }
    return std::make_pair(x, fx);
}

// ---------------------------------------------------------------------------------------------------
// TOMS 748 (Alefeld, Potra, Shi): enclosing root finder mixing secant, inverse-cubic and "newton-quadratic"
// steps with a forced bisection when the bracket shrinks by less than half.
// ---------------------------------------------------------------------------------------------------
namespace {

double guarded_div(double num, double den, double fallback)
{
    if (fabs(den) < 1 && fabs(den * DBL_MAX) <= fabs(num)) return fallback;
    return num / den;
}

double secant_point(double a, double b, double fa, double fb)
{
    const double tol = DBL_EPSILON * 5;
    const double c = a - (fa / (fb - fa)) * (b - a);
    if (c <= a + fabs(a) * tol || c >= b - fabs(b) * tol) return (a + b) / 2;
    return c;
}

double quadratic_point(double a, double b, double d, double fa, double fb, double fd, unsigned newton_steps)
{
    const double B = guarded_div(fb - fa, b - a, DBL_MAX);
    double A = guarded_div(fd - fb, d - b, DBL_MAX);
    A = guarded_div(A - B, d - a, 0.0);
    if (A == 0) return secant_point(a, b, fa, fb);
    double c = (sgn(A) * sgn(fa) > 0) ? a : b;
    for (unsigned i = 1; i <= newton_steps; ++i)
        c -= guarded_div(fa + (B + A * (c - b)) * (c - a), B + A * (2 * c - a - b), 1 + c - a);
    if (c <= a || c >= b) c = secant_point(a, b, fa, fb);
    return c;
}

double cubic_point(double a, double b, double d, double e, double fa, double fb, double fd, double fe)
{
    const double q11 = (d - e) * fd / (fe - fd);
    const double q21 = (b - d) * fb / (fd - fb);
    const double q31 = (a - b) * fa / (fb - fa);
    const double d21 = (b - d) * fd / (fd - fb);
    const double d31 = (a - b) * fb / (fb - fa);
    const double q22 = (d21 - q11) * fb / (fe - fb);
    const double q32 = (d31 - q21) * fa / (fd - fa);
    const double d32 = (d31 - q21) * fd / (fd - fa);
    const double q33 = (d32 - q22) * fa / (fe - fa);
    double c = q31 + q32 + q33 + a;
    if (c <= a || c >= b) c = quadratic_point(a, b, d, fa, fb, fd, 3);
    return c;
}

// Evaluate f at c (nudged off the ends), shrink [a,b] to the half that still brackets, remember the dropped end.
void rebracket(const std::function<double(double)>& f, double& a, double& b, double c, double& fa, double& fb,
               double& d, double& fd)
{
    const double tol = DBL_EPSILON * 2;
    if ((b - a) < 2 * tol * a)
        c = a + (b - a) / 2;
    else if (c <= a + fabs(a) * tol)
        c = a + fabs(a) * tol;
    else if (c >= b - fabs(b) * tol)
        c = b - fabs(b) * tol;
    const double fc = f(c);
    if (fc == 0) {
        a = c;
        fa = 0;
        d = 0;
        fd = 0;
        return;
    }
    if (sgn(fa) * sgn(fc) < 0) {
        d = b;
        fd = fb;
        b = c;
        fb = fc;
    } else {
        d = a;
        fd = fa;
        a = c;
        fa = fc;
    }
}

bool all_distinct(double fa, double fb, double fd, double fe)
{
    const double tiny = DBL_MIN * 32;
    return !(fabs(fa - fb) < tiny || fabs(fa - fd) < tiny || fabs(fa - fe) < tiny || fabs(fb - fd) < tiny ||
             fabs(fb - fe) < tiny || fabs(fd - fe) < tiny);
}

}  // namespace

// max_iter is IN-OUT exactly as in Boost: on entry the budget of function evaluations (two of which go to the end
// points), on return the number actually used.  The reference passes ONE variable to both of its calls (:780, :785-786),
// so its second root search may use only as many evaluations as the first one needed -- and returns a bracket that has
// not converged when that is too few.  The arithmetic is unsigned and wraps like the original.
std::pair<double, double> solve_toms748(const std::function<double(double)>& f, double ax, double bx, int tol_bits,
                                        std::uintmax_t& max_iter)
{
    max_iter -= 2;
    struct Restore {  // the four-argument Boost overload adds the two end-point evaluations back on every way out
        std::uintmax_t& m;
        ~Restore() { m += 2; }
    } restore{max_iter};
    double eps = ldexp(1.0, 1 - tol_bits);
    if (eps < 4 * DBL_EPSILON) eps = 4 * DBL_EPSILON;
    auto converged = [eps](double a, double b) { return fabs(a - b) <= eps * fmin(fabs(a), fabs(b)); };

    double a = ax, b = bx;
    double fa = f(ax), fb = f(bx);
    if (a >= b) throw std::domain_error("Error in function boost::math::tools::toms748_solve<double>: Parameters a and b out of order: a=" + std::to_string(a));
    if (converged(a, b) || fa == 0 || fb == 0) {
        max_iter = 0;
        if (fa == 0)
            b = a;
        else if (fb == 0)
            a = b;
        return std::make_pair(a, b);
    }
    if (sgn(fa) * sgn(fb) > 0) {
        char buf[64];
        snprintf(buf, sizeof(buf), "%.17g", a);
        throw std::domain_error(
            std::string("Error in function boost::math::tools::toms748_solve<double>: Parameters a and b do not "
                        "bracket the root: a=") + buf);
    }
    std::uintmax_t count = max_iter;
    double c, d = 1e5F, fd = 1e5F, e = 1e5F, fe = 1e5F;

    c = secant_point(a, b, fa, fb);
    rebracket(f, a, b, c, fa, fb, d, fd);
    --count;
    if (count && fa != 0 && !converged(a, b)) {
        c = quadratic_point(a, b, d, fa, fb, fd, 2);
        e = d;
        fe = fd;
        rebracket(f, a, b, c, fa, fb, d, fd);
        --count;
    }
    while (count && fa != 0 && !converged(a, b)) {
        const double a0 = a, b0 = b;
        c = all_distinct(fa, fb, fd, fe) ? cubic_point(a, b, d, e, fa, fb, fd, fe)
                                         : quadratic_point(a, b, d, fa, fb, fd, 2);
        e = d;
        fe = fd;
        rebracket(f, a, b, c, fa, fb, d, fd);
        if (0 == --count || fa == 0 || converged(a, b)) break;

        c = all_distinct(fa, fb, fd, fe) ? cubic_point(a, b, d, e, fa, fb, fd, fe)
                                         : quadratic_point(a, b, d, fa, fb, fd, 3);
        rebracket(f, a, b, c, fa, fb, d, fd);
        if (0 == --count || fa == 0 || converged(a, b)) break;

        // double-length secant step from the end with the smaller |f|
        double u, fu;
        if (fabs(fa) < fabs(fb)) {
            u = a;
            fu = fa;
        } else {
            u = b;
            fu = fb;
        }
        c = u - 2 * (fu / (fb - fa)) * (b - a);
        if (fabs(c - u) > (b - a) / 2) c = a + (b - a) / 2;
        e = d;
        fe = fd;
        rebracket(f, a, b, c, fa, fb, d, fd);
        if (0 == --count || fa == 0 || converged(a, b)) break;

        if ((b - a) < 0.5 * (b0 - a0)) continue;
        e = d;
        fe = fd;
        rebracket(f, a, b, a + (b - a) / 2, fa, fb, d, fd);
        --count;
    }
    max_iter -= count;
    if (fa == 0)
        b = a;
    else if (fb == 0)
        a = b;
    return std::make_pair(a, b);
}

// ---------------------------------------------------------------------------------------------------
// fit helpers (ACFcalculator.cpp:418-511)
// ---------------------------------------------------------------------------------------------------
namespace {

// least squares over [lower, upper): slope m, intercept b, r2 = 1 - SS_res / sum(y^2)   (:418-442)
void least_squares(const std::vector<double>& x, const std::vector<double>& y, int lower, int upper, double& m,
                   double& b, double& r2)
{
    const int n = upper - lower;
    const double sx = std::accumulate(x.begin() + lower, x.begin() + upper, 0.0);
    const double sy = std::accumulate(y.begin() + lower, y.begin() + upper, 0.0);
    const double sxx = std::inner_product(x.begin() + lower, x.begin() + upper, x.begin() + lower, 0.0);
    const double sxy = std::inner_product(x.begin() + lower, x.begin() + upper, y.begin() + lower, 0.0);
    const double syy = std::inner_product(y.begin() + lower, y.begin() + upper, y.begin() + lower, 0.0);
    m = (n * sxy - sx * sy) / (n * sxx - sx * sx);
    b = (sy - m * sx) / n;
    double res_sum = 0.0;
    for (int i = lower; i < upper; ++i) {
        const double res = y.at(i) - (m * x.at(i) + b);
        res_sum += res * res;
    }
    r2 = 1.0 - res_sum / syy;
}

// forward differences with an extrapolated first entry (:463-478)
std::vector<double> slopes(const std::vector<double>& x, const std::vector<double>& y)
{
    std::vector<double> df;
    const double s1 = (y.at(1) - y.at(0)) / (x.at(1) - x.at(0));
    const double s2 = (y.at(2) - y.at(1)) / (x.at(2) - x.at(1));
    df.push_back(s1 - (s2 - s1));
    for (unsigned int i = 0; i < y.size() - 1; ++i) df.push_back((y.at(i + 1) - y.at(i)) / (x.at(i + 1) - x.at(i)));
    return df;
}

// index of the entry with the smallest magnitude (:481-495)
int argmin_abs(const std::vector<double>& y)
{
    int best = 0;
    double best_val = fabs(y.at(0));
    for (unsigned int i = 0; i < y.size(); ++i)
        if (fabs(y.at(i)) < best_val) {
            best = i;
            best_val = fabs(y.at(i));
        }
    return best;
}

// 5-point running mean with the reference's end treatment, including its swapped last two entries (:498-511)
std::vector<double> running_mean5(const std::vector<double>& y)
{
    std::vector<double> out;
    out.push_back((y.at(0) + y.at(1) + y.at(2)) / 3.0);
    out.push_back((y.at(0) + y.at(1) + y.at(2) + y.at(3)) / 4.0);
    for (unsigned int i = 2; i < y.size() - 2; ++i)
        out.push_back((y.at(i - 2) + y.at(i - 1) + y.at(i) + y.at(i + 1) + y.at(i + 2)) / 5.0);
    out.push_back(y.at(y.size() - 1));
    out.push_back(y.at(y.size() - 2));
    return out;
}

}  // namespace

GleResult run_gle_stage(const GleInput& in) { return run_gle_stage(in, std::cout); }

GleResult run_gle_stage(const GleInput& in, std::ostream& log)
{
    const double* vacf = in.vacf;
    const int ncorr = in.ncorr;
    const double var = in.var, var_vel = in.var_vel, dt = in.timestep;

    // denominator of D(s) and D(s) itself (ACFcalculator.cpp:90-103)
    auto denom = [&](double s) {
        const double lv = laplace(vacf, dt, s, ncorr);
        return lv * (s * var + var_vel / s) - var * var_vel;
    };
    auto ds_of = [&](double s) {
        const double lv = laplace(vacf, dt, s, ncorr);
        return -(lv * var * var_vel) / (lv * (s * var + var_vel / s) - var * var_vel);
    };
    // D(s) sampled from s to s_max, positive values only (:514-530).  The abscissae come from the reference's own
    // recurrence s = s + s_delta; the evaluations are independent of each other, so they are spread over host threads
    // (each value is computed exactly as in the serial loop, the results are collected in order: same bits).
    auto sample_ds = [&](double s, double s_max, double s_delta, std::vector<double>& sv, std::vector<double>& dv) {
        std::vector<double> grid;
        while (s <= s_max) {
            grid.push_back(s);
            s = s + s_delta;
        }
        std::vector<double> val(grid.size());
        auto evaluate = [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                const double lv = laplace(vacf, dt, grid[i], ncorr);
                const double den = lv * (grid[i] * var + var_vel / grid[i]) - var * var_vel;
                val[i] = -(lv * var * var_vel) / den;
            }
        };
        size_t nthreads = in.threads > 0 ? (size_t)in.threads : std::thread::hardware_concurrency();
        const double work = (double)grid.size() * (double)ncorr;  // exp() evaluations
        if (nthreads > grid.size() / 8) nthreads = grid.size() / 8;
        if (nthreads < 2 || work < 2e5) {
            evaluate(0, grid.size());
        } else {
            std::vector<std::thread> pool;
            const size_t per = (grid.size() + nthreads - 1) / nthreads;
            for (size_t t = 1; t < nthreads; ++t) {
                const size_t lo = t * per, hi = lo + per < grid.size() ? lo + per : grid.size();
                if (lo < hi) pool.emplace_back(evaluate, lo, hi);
            }
            evaluate(0, per < grid.size() ? per : grid.size());
            for (auto& th : pool) th.join();
        }
        for (size_t i = 0; i < grid.size(); ++i)
            if (val[i] > 0.0) {
                sv.push_back(grid[i]);
                dv.push_back(val[i]);
            }
    };

    GleResult out;
    const int bits = DBL_MANT_DIG;  // std::numeric_limits<double>::digits (:546)
    // Step 1 (:778-814): the two singularities either side of the denominator's minimum
    const std::pair<double, double> den_min = minimise_brent(denom, kSStart, kSEnd, bits);
    std::uintmax_t max_iter = 500;  // one variable for both searches, as upstream (:780): see solve_toms748
    const std::pair<double, double> root1 = solve_toms748(denom, kSStart, den_min.first, 30, max_iter);
    const std::pair<double, double> root2 = solve_toms748(denom, den_min.first, kSEnd, 30, max_iter);
    const double den_tol = den_min.second / 100;
    double s1 = root1.first;
    while (denom(s1) > den_tol) s1 += kSIncrement;
    double s2 = root2.first;
    while (denom(s2) > den_tol) s2 -= kSIncrement;
    const std::pair<double, double> ds_min = minimise_brent(ds_of, s1, s2, bits);
    log << "#s1 " << s1 << std::endl;
    log << "#s2 " << s2 << std::endl;
    log << "#Ds_min " << ds_min.second << std::endl;
    out.s1 = s1;
    out.s2 = s2;
    out.ds_min = ds_min.second;

    // Step 2 (:816-830)
    const double s_range = s2 - ds_min.first;
    if (s_range <= 0.0) throw std::out_of_range("No minimum found within singularities");
    std::vector<double> s_values, ds_values;
    sample_ds(kSIncrement, s2, s_range / kFitWidth1, s_values, ds_values);

    // Step 3 (:832-888): second derivative, smoothed; widest low-curvature stretch around its minimum
    const std::vector<double> df = slopes(s_values, ds_values);
    (void)running_mean5(df);  // the reference smooths df too (:838) and never uses it; kept for its range checks
    const std::vector<double> ddf = slopes(s_values, df);
    const std::vector<double> ddf_smooth = running_mean5(ddf);
    std::vector<double> s_pos, dd_pos;
    for (unsigned int i = 0; i < ddf_smooth.size(); ++i)
        if (ddf_smooth.at(i) > 0.0) {
            s_pos.push_back(s_values.at(i));
            dd_pos.push_back(ddf_smooth.at(i));
        }
    const int imin = argmin_abs(dd_pos);
    int lower = imin - 1;
    out.dd_ds_min = dd_pos.at(imin);
    while (lower > 0 && dd_pos.at(lower) < dd_pos.at(imin) * 2.0) --lower;
    int upper = imin + 1;
    while (dd_pos.at(upper) < dd_pos.at(imin) * 3.0 && (size_t)upper < dd_pos.size() - 1) ++upper;
    while ((upper - lower) < kSegMin * kFitWidth1) {
        if (dd_pos.at(upper) < dd_pos.at(lower) && lower > 0)
            --lower;
        else if ((size_t)upper < dd_pos.size())
            ++upper;
        else
            break;
    }
    const double fit_lo = s_pos.at(lower);
    const double fit_hi = s_pos.at(upper);
    std::vector<double> s_fit, ds_fit;
    sample_ds(fit_lo, fit_hi, (fit_hi - fit_lo) / kFitWidth2, s_fit, ds_fit);

    // Step 4 (:896-899)
    least_squares(s_fit, ds_fit, 0, (int)s_fit.size(), out.m, out.b, out.r2);
    return out;
}

}  // namespace acfhost
